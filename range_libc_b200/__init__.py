"""range_libc_b200 -- B200-native backend for range_libc's batched 2-D ray casting path.

Host-side mirror of the reference's Python module ``range_libc``
(/root/reference/pywrapper/RangeLibc.pyx): the same class and method names
(PyOMap, PyBresenhamsLine, PyRayMarching, PyRayMarchingGPU, PyCDDTCast, PyGiantLUTCast; calc_range,
calc_range_many, calc_range_repeat_angles, calc_range_repeat_angles_eval_sensor_model,
eval_sensor_model, set_sensor_model, prune), forwarding to the C ABI in
include/rangelib_b200.h.  Arrays may be numpy arrays (host; blocking, like the reference) or
torch CUDA tensors (device; asynchronous on the method's stream).

Every range method here runs on the GPU.  There is no CPU path: importing works anywhere, but
constructing a method without the compiled library or without a B200 raises.
"""
from .api import (PyBresenhamsLine, PyCDDTCast, PyGiantLUTCast, PyOMap, PyRayMarching, PyRayMarchingGPU,  # noqa: F401
                  kernel_launches)
from .cabi import RangeLibError  # noqa: F401

__all__ = ["PyOMap", "PyBresenhamsLine", "PyRayMarching", "PyRayMarchingGPU", "PyCDDTCast", "PyGiantLUTCast", "RangeLibError",
           "kernel_launches"]
