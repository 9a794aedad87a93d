"""Model package of the drop-in: only the NCSN++ family of the hot path is provided
(the reference's models/__init__.py also pulls in unet1d / adm / vdm, which are out of scope)."""
from . import utils, layers, layerspp, ncsnpp  # noqa: F401
