"""Stand-in for the one pytorch3d function the reference imports (layers/utils.py:6), for machines where pytorch3d is
not installed.  Put `dropin/_optional_shims` on sys.path ONLY in that case; a real pytorch3d is used untouched."""
