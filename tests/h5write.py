"""Shim: the HDF5 writer lives in the package (bench.py feeds CnnBuilder through it as well)."""
from crcnn_b200.h5write import *  # noqa: F401,F403
from crcnn_b200.h5write import write_h5  # noqa: F401
