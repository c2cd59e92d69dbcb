/* Stand-in for the STOCK libbifrost.so in tests/test_overlay.py: it exports
 *   - the symbols of SURVEY 8(b)'s list that libbifrost_b200.so does not have
 *     (ring, proclog, affinity: src/bifrost/ring.h:74-227, proclog.h,
 *     affinity.h) as minimal in-memory implementations, and
 *   - the hot-path symbols too, every one of which reports how often it was
 *     called and returns BF_STATUS_UNSUPPORTED,
 * so that the test can tell which library served a call after the overlay.
 * TEST INFRASTRUCTURE ONLY: nothing here is product code. */
#include <stdlib.h>
#include <string.h>

typedef int BFstatus;
enum { BF_STATUS_SUCCESS = 0, BF_STATUS_UNSUPPORTED = 7 };

static int g_hot_calls = 0;
int standin_hot_calls(void) { return g_hot_calls; }

#define HOT(name) BFstatus name(void) { ++g_hot_calls; return BF_STATUS_UNSUPPORTED; }
HOT(bfTranspose) HOT(bfReduce) HOT(bfUnpack) HOT(bfQuantize)
HOT(bfFdmtCreate) HOT(bfFdmtInit) HOT(bfFdmtSetStream) HOT(bfFdmtExecute) HOT(bfFdmtDestroy)
HOT(bfFftCreate) HOT(bfFftInit) HOT(bfFftExecute) HOT(bfFftDestroy)
HOT(bfLinAlgCreate) HOT(bfLinAlgDestroy) HOT(bfLinAlgMatMul)

/* ---- ring (a name, a space and a byte count are all the test looks at) */
typedef struct { char name[64]; int space; unsigned long long nbyte; } Ring;
BFstatus bfRingCreate(void** ring, const char* name, int space) {
	Ring* r = (Ring*)calloc(1, sizeof(Ring));
	if( !r ) return 1;
	strncpy(r->name, name ? name : "", sizeof(r->name) - 1);
	r->space = space;
	*ring = r;
	return BF_STATUS_SUCCESS;
}
BFstatus bfRingDestroy(void* ring) { free(ring); return BF_STATUS_SUCCESS; }
BFstatus bfRingResize(void* ring, unsigned long long span, unsigned long long total, unsigned long long nringlet) {
	(void)span; (void)nringlet;
	((Ring*)ring)->nbyte = total;
	return BF_STATUS_SUCCESS;
}
BFstatus bfRingGetName(void* ring, const char** name) { *name = ((Ring*)ring)->name; return BF_STATUS_SUCCESS; }
BFstatus bfRingGetSpace(void* ring, int* space) { *space = ((Ring*)ring)->space; return BF_STATUS_SUCCESS; }

/* ---- proclog / affinity */
static int g_core = -1;
BFstatus bfProcLogCreate(void** log, const char* name) { (void)name; *log = malloc(1); return BF_STATUS_SUCCESS; }
BFstatus bfProcLogDestroy(void* log) { free(log); return BF_STATUS_SUCCESS; }
BFstatus bfProcLogUpdate(void* log, const char* contents) { (void)log; (void)contents; return BF_STATUS_SUCCESS; }
BFstatus bfAffinityGetCore(int* core) { *core = g_core; return BF_STATUS_SUCCESS; }
BFstatus bfAffinitySetCore(int core) { g_core = core; return BF_STATUS_SUCCESS; }

/* ---- stream / device: the stock side of the mirrored setters */
static int g_stream_sets = 0, g_device = -1;
int standin_stream_sets(void) { return g_stream_sets; }
int standin_device(void) { return g_device; }
BFstatus bfStreamSet(void const* stream) { (void)stream; ++g_stream_sets; return BF_STATUS_SUCCESS; }
BFstatus bfDeviceSet(int device) { g_device = device; return BF_STATUS_SUCCESS; }
