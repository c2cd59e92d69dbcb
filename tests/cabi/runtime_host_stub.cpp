// Host-only stand-ins for the few runtime entry points csrc/ring.cpp calls
// (csrc/runtime.cu proper needs the CUDA runtime): lets the ring be built with
// -fsanitize=thread and stressed without a device (tests/test_ring_stress.py).
// TEST INFRASTRUCTURE.
#include <bifrost_b200.h>
#include <cstdlib>
#include <cstring>

namespace bfb { void report_failure(const char*, const char*, int, BFstatus) {} }

extern "C" {
const char* bfGetStatusString(BFstatus) { return "status"; }
const char* bfGetSpaceString(BFspace)   { return "system"; }
BFsize      bfGetAlignment(void)        { return 4096; }
BFstatus bfMalloc(void** ptr, BFsize size, BFspace) {
	return ::posix_memalign(ptr, 4096, size ? size : 1) ? BF_STATUS_MEM_ALLOC_FAILED : BF_STATUS_SUCCESS;
}
BFstatus bfFree(void* ptr, BFspace) { ::free(ptr); return BF_STATUS_SUCCESS; }
BFstatus bfMemcpy2D(void* dst, BFsize dst_stride, BFspace, const void* src, BFsize src_stride, BFspace,
                    BFsize width, BFsize height) {
	for( BFsize r=0; r<height; ++r ) ::memcpy((char*)dst + r*dst_stride, (const char*)src + r*src_stride, width);
	return BF_STATUS_SUCCESS;
}
BFstatus bfStreamSynchronize(void) { return BF_STATUS_SUCCESS; }
}
