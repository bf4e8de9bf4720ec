import io
import os

SYNTHETIC_MAGIC = b'HFAGP_SYNTHETIC_GENERATOR'


def open_url(url, *args, **kwargs):
    """Context manager yielding a binary file, like ``dnnlib.util.open_url`` for local paths.
    With ``HFAGP_SYNTHETIC_GENERATOR=1`` and no file at ``url`` it yields a marker stream that
    ``legacy.load_network_pkl`` turns into a seeded random-init generator (no EG3D pickle exists offline)."""
    if os.path.isfile(url):
        return open(url, 'rb')
    if os.environ.get('HFAGP_SYNTHETIC_GENERATOR') == '1':
        return io.BytesIO(SYNTHETIC_MAGIC)
    raise FileNotFoundError(url)
