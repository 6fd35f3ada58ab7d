"""Drop-in replacement for the reference's chamfer_extension/dist_chamfer.py (no compiled `chamfer` extension needed):
`from chamfer_extension import dist_chamfer` in train_inpainter.py:11 resolves here when dropin/ is ahead of the
reference checkout on sys.path."""
from cloud_transformers_b200.chamfer import (  # noqa: F401
    ChamferFunction, ChamferDist, loss_chamfer, loss_chamfer_adj, loss_chamder_2d)
