"""Stand-in for the reference's ``cpnp`` extension module (binary only, absent from the
reference tree; call sites val.py:200-202): ``cpnp.cpnp`` / ``cpnp.cpnp_m`` on the CUDA LM."""
from .pnp import cpnp, cpnp_m  # noqa: F401
