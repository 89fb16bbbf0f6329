"""demo/utils.py of the reference: the ASCII PLY loader (re-exported from pc_sam.utils.ply)."""
from pc_sam.utils.ply import load_ply  # noqa: F401
