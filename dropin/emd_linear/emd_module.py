"""Drop-in for the reference's emd_linear/emd_module.py: same names (emdFunction, emdModule), same call, running on
libctb200's one-launch auction kernel instead of the `emd` CUDA extension.  With dropin/ ahead of a reference checkout
on sys.path, `import emd_linear.emd_module as emd` (train_inpainter.py:12) resolves here -- `emd_linear` has no
__init__.py in the reference, so it is a namespace package and only this module is overridden."""
from cloud_transformers_b200.emd import emdFunction, emdModule  # noqa: F401
