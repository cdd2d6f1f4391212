"""motionestimation_b200 -- B200-native full-search block-matching motion estimation.

Thin ctypes binding over the C ABI in ``include/me_b200.h`` (built in-tree as
``motionestimation_b200/libme_b200.so`` by ``make -C motionestimation_b200``).
The product is the CUDA library + the plain-C host layer; this module only
loads it so Python tests and ``bench.py`` can call the same entry points a C
caller links against.  There is no Python or CPU implementation of the search
here: if the library is missing, importing :mod:`motionestimation_b200.lib`
symbols raises, and every compute call returns ``ME_ERR_NO_DEVICE`` without a GPU.

Reference path being replaced: ``src/cpu/main.c:18-107,144-158`` of
souravBhat/MotionEstimation.
"""
from .lib import (  # noqa: F401
    ME_OK, ME_ERR_INVALID_ARG, ME_ERR_UNSUPPORTED, ME_ERR_CUDA, ME_ERR_NO_DEVICE,
    ME_ERR_NOMEM, ME_ERR_STATE, ME_KERNEL_AUTO, ME_KERNEL_GENERIC, ME_KERNEL_TILED, ME_KERNEL_DIRECT,
    ME_KERNEL_SSIM, ME_KERNEL_FAST, ME_HOST_WRITE_COMBINED,
    ME_COST_MSE, ME_COST_SSIM, ME_SEARCH_FULL, ME_SEARCH_THREE_STEP, ME_SEARCH_DIAMOND,
    MeError, Block, PredictionFrame, Field, load_library, library_path, device_count,
    Estimator, create_prediction_frame, search_prediction_frame, int_peak, PEAK_NAMES,
)
from .frames import (  # noqa: F401
    read_yuv_luma, tiled_frames, shifted_noise_pair, constant_pair, random_pair, checker_pair, far_pair, inverted_pair, foreman,
    block_grid, pixel_compares, candidates,
)

__all__ = [n for n in dir() if not n.startswith("_")]
