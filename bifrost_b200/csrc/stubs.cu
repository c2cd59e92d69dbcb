// stubs.cu -- entry points declared in bifrost_b200.h whose kernels have not
// landed yet.  Each returns BF_STATUS_UNSUPPORTED (never a CPU fallback) and
// is removed from this file when its real implementation is added.
#include "core.hpp"

extern "C" {

#if 0
BFstatus bfUnpack(BFarray const*, BFarray const*, BFbool) { return BF_STATUS_UNSUPPORTED; }
#endif

#if 0
BFstatus bfFftCreate(BFfft* plan) { if( plan ) *plan = nullptr; return BF_STATUS_UNSUPPORTED; }
BFstatus bfFftInit(BFfft, BFarray const*, BFarray const*, int, int const*, BFbool, size_t*) { return BF_STATUS_UNSUPPORTED; }
BFstatus bfFftExecute(BFfft, BFarray const*, BFarray const*, BFbool, void*, size_t) { return BF_STATUS_UNSUPPORTED; }
BFstatus bfFftDestroy(BFfft) { return BF_STATUS_UNSUPPORTED; }
#endif
#if 0
BFstatus bfSpectrometerFused(BFarray const*, BFarray const*, int, int, double) { return BF_STATUS_UNSUPPORTED; }
#endif

#if 0
BFstatus bfLinAlgCreate(BFlinalg* h) { if( h ) *h = nullptr; return BF_STATUS_UNSUPPORTED; }
BFstatus bfLinAlgDestroy(BFlinalg) { return BF_STATUS_UNSUPPORTED; }
BFstatus bfLinAlgMatMul(BFlinalg, double, BFarray const*, BFarray const*, double, BFarray const*) { return BF_STATUS_UNSUPPORTED; }
#endif

} // extern "C"
