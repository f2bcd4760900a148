"""jpgb200 -- B200-native (sm_100a) JPEG encode hot path behind the C API of
matthewT53/JPEG-Encoder-and-Decoder.

The product is csrc/ (CUDA kernels + the extern "C" boundary of include/jpgb200.h) and host/
(the C mirror of the reference's encoder/BMP API).  This Python package only binds the C ABI
for tests/ and bench.py.  The directory name contains hyphens, so load it through
`__graft_entry__.load_package()` (importlib) under the module name `jpgb200`.
"""
from . import binding, build, mgpu  # noqa: F401
from .binding import Encoder, Jpgb200Error  # noqa: F401
