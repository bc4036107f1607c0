"""fov360-b200: B200-native (sm_100a CUDA) foveation transform behind the reference's
SATEncoder / SATDecoder / ImageSampler interfaces.

The product is ``libfov360.so`` (csrc/, C ABI in include/fov360.h).  This Python package only
builds and binds it for the tests and the benchmark; the directory name contains hyphens, so load
it with ``importlib.import_module("foveated-360-video_b200")`` (see ``__graft_entry__.py``).
"""
from . import build as build_module  # noqa: F401
from . import sharding  # noqa: F401
from ._capi import PROTOTYPES, header_symbols, library_path, load  # noqa: F401
from .host import (  # noqa: F401
    DeviceBuffer,
    EncodeSampleFramesGPU,
    FoveateFramesDeviceGazeGPU,
    FoveateFramesGPU,
    FovError,
    GazeViewPoints,
    ImageSampler,
    OpenCLManager,
    Projections,
    SATDecoder,
    SATEncoder,
    VideoFrameConverter,
    reduced_dim,
)

REDUCED_BUFFER_WIDTH = 1072   # parameters.h:8
REDUCED_BUFFER_HEIGHT = 608   # parameters.h:9
