"""Device / stream control (mirrors python/bifrost/device.py)."""
import ctypes
from bifrost_b200.libbifrost import _bf, _check, _get


def set_device(device):
    if isinstance(device, int):
        _check(_bf.bfDeviceSet(device))
    else:
        _check(_bf.bfDeviceSetById(str(device).encode()))


def get_device():
    return _get(_bf.bfDeviceGet)


def set_devices_no_spin_cpu():
    _check(_bf.bfDevicesSetNoSpinCPU())


def stream_synchronize():
    _check(_bf.bfStreamSynchronize())


def get_stream():
    """The calling thread's CUDA stream handle as an int."""
    handle = ctypes.c_void_p()
    _check(_bf.bfStreamGet(ctypes.byref(handle)))
    return handle.value or 0


def set_stream(stream):
    """Make `stream` (int handle or object with .cuda_stream) this thread's stream."""
    handle = ctypes.c_void_p(int(getattr(stream, 'cuda_stream', stream)))
    _check(_bf.bfStreamSet(ctypes.byref(handle)))


class ExternalStream(object):
    """Context manager: run bifrost_b200 calls on an externally owned stream
    (e.g. ``torch.cuda.current_stream().cuda_stream``)."""

    def __init__(self, stream):
        self._new = int(getattr(stream, 'cuda_stream', stream))

    def __enter__(self):
        self._old = get_stream()
        set_stream(self._new)
        return self

    def __exit__(self, *exc):
        set_stream(self._old)
