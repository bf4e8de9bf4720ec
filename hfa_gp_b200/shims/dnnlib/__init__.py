"""Stand-in for NVlabs/eg3d's ``dnnlib`` — only what ``code/networks/headnerf.py:32`` touches
(``dnnlib.util.open_url``).  Put ``hfa_gp_b200/shims`` on ``sys.path`` to run the unmodified reference
``headnerf.py`` on the B200 generator."""
from . import util  # noqa: F401
