"""Times the fused spectrometer gulp (config 3: 32 frames x 4096 chan x 4096 fine_time x 2 pol ci8)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bifrost_b200 as bf
from bifrost_b200.libbifrost import _bf, _check
from tools.bench_ops import timeit
stream = torch.cuda.current_stream(); bf.device.set_stream(stream.cuda_stream)
nframe, nchan, nfft = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 4096, 4096
x = bf.empty((nframe, nchan, nfft, 2), "ci8", "cuda")
raw = torch.randint(-127, 128, (nframe * nchan * nfft * 4,), dtype=torch.int8, device="cuda")
_check(_bf.bfMemcpy(x.ctypes.data, 2, raw.data_ptr(), 2, raw.numel()))
o = bf.zeros((4, nchan * nfft // 4), "f32", "cuda")
ms = timeit(lambda: bf.spectrometer(x, o, 4096, 4, 0.0), nrep=5, stream=stream)
print("fused spectrometer %d frames: ms %.3f  Msamples/s %.0f  in GB/s %.0f" % (
    nframe, ms, nframe * nchan * nfft / ms / 1e3, nframe * nchan * nfft * 4 / ms / 1e6))
