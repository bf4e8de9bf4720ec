"""TF32 dense matmul peak on this GPU (BASELINE.md §2 / VERDICT r1 item 6): torch.matmul fp32 inputs with TF32 tensor cores,
8192^3, best of 10 (burst) and back to back for ~3 s (sustained); bf16 beside it for the ratio."""
import time, json, torch
dev = 'cuda'
n = 8192
out = {}
for name, dt, tf32 in (('tf32', torch.float32, True), ('bf16', torch.bfloat16, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device=dev, dtype=dt); b = torch.randn(n, n, device=dev, dtype=dt)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 0; t0 = time.time(); e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(20):
            a @ b
        reps += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    out[name] = {'burst_tflops': 2 * n ** 3 / best / 1e9, 'sustained_tflops': 2 * n ** 3 * reps / e0.elapsed_time(e1) / 1e9}
print(json.dumps(out))
