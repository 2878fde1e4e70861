"""Numpy stand-in for the `tensorflow` package -- TEST INFRASTRUCTURE ONLY.

TensorFlow is not installable in this image (no network).  The reference
(`/root/reference/Main_Functions.py:3-4`) does `import tensorflow.compat.v1 as tf`
and only uses ~25 eager-expressible ops from it; `compat/v1.py` implements those
with numpy so the reference module can be imported and executed UNMODIFIED.
Nothing outside `oracle/` and `tests/` may import this package.
"""
