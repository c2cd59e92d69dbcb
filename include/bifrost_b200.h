/* bifrost_b200.h -- C ABI of the B200-native Bifrost hot path.
 *
 * One consolidated header for the drop-in boundary.  Every type, enum value
 * and entry point below is what the reference's ctypes FFI binds for the
 * per-gulp DSP path; names, argument order and status codes are kept so that
 * `libbifrost_b200.so` can be loaded in place of `libbifrost.so` for this
 * path.  Each declaration cites the reference interface it replaces
 * (paths relative to the reference tree).
 *
 * The per-topic headers under include/bifrost/ (fdmt.h, fft.h, ...) simply
 * include this file, so C/C++ callers that `#include <bifrost/fdmt.h>` keep
 * compiling.
 */
#ifndef BIFROST_B200_H_
#define BIFROST_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ *
 * Scalars and status codes            (ref: src/bifrost/common.h:40-84)
 * ------------------------------------------------------------------ */
typedef int                BFbool;
typedef float              BFcomplex[2];
typedef float              BFreal;
typedef unsigned long      BFsize;
typedef unsigned long long BFoffset;
typedef signed long long   BFdelta;

typedef enum BFstatus_ {
	BF_STATUS_SUCCESS              = 0,
	BF_STATUS_END_OF_DATA          = 1,
	BF_STATUS_WOULD_BLOCK          = 2,
	BF_STATUS_INVALID_POINTER      = 8,
	BF_STATUS_INVALID_HANDLE       = 9,
	BF_STATUS_INVALID_ARGUMENT     = 10,
	BF_STATUS_INVALID_STATE        = 11,
	BF_STATUS_INVALID_SPACE        = 12,
	BF_STATUS_INVALID_SHAPE        = 13,
	BF_STATUS_INVALID_STRIDE       = 14,
	BF_STATUS_INVALID_DTYPE        = 15,
	BF_STATUS_MEM_ALLOC_FAILED     = 32,
	BF_STATUS_MEM_OP_FAILED        = 33,
	BF_STATUS_UNSUPPORTED          = 48,
	BF_STATUS_UNSUPPORTED_SPACE    = 49,
	BF_STATUS_UNSUPPORTED_SHAPE    = 50,
	BF_STATUS_UNSUPPORTED_STRIDE   = 51,
	BF_STATUS_UNSUPPORTED_DTYPE    = 52,
	BF_STATUS_FAILED_TO_CONVERGE   = 64,
	BF_STATUS_INSUFFICIENT_STORAGE = 65,
	BF_STATUS_DEVICE_ERROR         = 66,
	BF_STATUS_INTERNAL_ERROR       = 99
} BFstatus;

const char* bfGetStatusString(BFstatus status);   /* common.h:81 */
BFbool      bfGetDebugEnabled(void);              /* common.h:82 */
BFstatus    bfSetDebugEnabled(BFbool enabled);    /* common.h:83 */
BFbool      bfGetCudaEnabled(void);               /* common.h:84 */

/* ------------------------------------------------------------------ *
 * Memory spaces                       (ref: src/bifrost/memory.h:43-82)
 * ------------------------------------------------------------------ */
typedef enum BFspace_ {
	BF_SPACE_AUTO         = 0,
	BF_SPACE_SYSTEM       = 1,
	BF_SPACE_CUDA         = 2,
	BF_SPACE_CUDA_HOST    = 3,
	BF_SPACE_CUDA_MANAGED = 4
} BFspace;

BFstatus    bfMalloc(void** ptr, BFsize size, BFspace space);
BFstatus    bfFree(void* ptr, BFspace space);
BFstatus    bfGetSpace(const void* ptr, BFspace* space);
const char* bfGetSpaceString(BFspace space);
/* Synchronous w.r.t. the host, asynchronous w.r.t. the device (memory.h:59) */
BFstatus    bfMemcpy(void* dst, BFspace dst_space,
                     const void* src, BFspace src_space, BFsize count);
BFstatus    bfMemcpy2D(void* dst, BFsize dst_stride, BFspace dst_space,
                       const void* src, BFsize src_stride, BFspace src_space,
                       BFsize width, BFsize height);
BFstatus    bfMemset(void* ptr, BFspace space, int value, BFsize count);
BFstatus    bfMemset2D(void* ptr, BFsize stride, BFspace space, int value,
                       BFsize width, BFsize height);
BFsize      bfGetAlignment(void);

/* ------------------------------------------------------------------ *
 * Array descriptor                    (ref: src/bifrost/array.h:38-234)
 * The dtype word is a bit-field: low byte = bits per real component,
 * 0xF00 = kind, 0xFF000 = vector length - 1, 0x100000 = complex.
 * sizeof(BFarray) == 168 and the field order is ABI.
 * ------------------------------------------------------------------ */
enum { BF_MAX_DIMS = 8 };

typedef enum BFdtype_ {
	BF_DTYPE_NBIT_BITS    = 0x0000FF,
	BF_DTYPE_TYPE_BITS    = 0x000F00,
	BF_DTYPE_VECTOR_BITS  = 0x0FF000,
	BF_DTYPE_VECTOR_BIT0  = 12,
	BF_DTYPE_COMPLEX_BIT  = 0x100000,

	BF_DTYPE_INT_TYPE     = 0x0000,
	BF_DTYPE_UINT_TYPE    = 0x0100,
	BF_DTYPE_FLOAT_TYPE   = 0x0200,
	BF_DTYPE_STRING_TYPE  = 0x0300,
	BF_DTYPE_STORAGE_TYPE = 0x0400,

	BF_DTYPE_I1   =  1 | BF_DTYPE_INT_TYPE,
	BF_DTYPE_I2   =  2 | BF_DTYPE_INT_TYPE,
	BF_DTYPE_I4   =  4 | BF_DTYPE_INT_TYPE,
	BF_DTYPE_I8   =  8 | BF_DTYPE_INT_TYPE,
	BF_DTYPE_I16  = 16 | BF_DTYPE_INT_TYPE,
	BF_DTYPE_I32  = 32 | BF_DTYPE_INT_TYPE,
	BF_DTYPE_I64  = 64 | BF_DTYPE_INT_TYPE,

	BF_DTYPE_U1   =  1 | BF_DTYPE_UINT_TYPE,
	BF_DTYPE_U2   =  2 | BF_DTYPE_UINT_TYPE,
	BF_DTYPE_U4   =  4 | BF_DTYPE_UINT_TYPE,
	BF_DTYPE_U8   =  8 | BF_DTYPE_UINT_TYPE,
	BF_DTYPE_U16  = 16 | BF_DTYPE_UINT_TYPE,
	BF_DTYPE_U32  = 32 | BF_DTYPE_UINT_TYPE,
	BF_DTYPE_U64  = 64 | BF_DTYPE_UINT_TYPE,

	BF_DTYPE_F16  = 16 | BF_DTYPE_FLOAT_TYPE,
	BF_DTYPE_F32  = 32 | BF_DTYPE_FLOAT_TYPE,
	BF_DTYPE_F64  = 64 | BF_DTYPE_FLOAT_TYPE,

	BF_DTYPE_CI1  =  1 | BF_DTYPE_INT_TYPE | BF_DTYPE_COMPLEX_BIT,
	BF_DTYPE_CI2  =  2 | BF_DTYPE_INT_TYPE | BF_DTYPE_COMPLEX_BIT,
	BF_DTYPE_CI4  =  4 | BF_DTYPE_INT_TYPE | BF_DTYPE_COMPLEX_BIT,
	BF_DTYPE_CI8  =  8 | BF_DTYPE_INT_TYPE | BF_DTYPE_COMPLEX_BIT,
	BF_DTYPE_CI16 = 16 | BF_DTYPE_INT_TYPE | BF_DTYPE_COMPLEX_BIT,
	BF_DTYPE_CI32 = 32 | BF_DTYPE_INT_TYPE | BF_DTYPE_COMPLEX_BIT,
	BF_DTYPE_CI64 = 64 | BF_DTYPE_INT_TYPE | BF_DTYPE_COMPLEX_BIT,

	BF_DTYPE_CF16 = 16 | BF_DTYPE_FLOAT_TYPE | BF_DTYPE_COMPLEX_BIT,
	BF_DTYPE_CF32 = 32 | BF_DTYPE_FLOAT_TYPE | BF_DTYPE_COMPLEX_BIT,
	BF_DTYPE_CF64 = 64 | BF_DTYPE_FLOAT_TYPE | BF_DTYPE_COMPLEX_BIT
} BFdtype;

typedef struct BFarray_ {
	void*   data;
	BFspace space;
	BFdtype dtype;
	int     ndim;
	long    shape[BF_MAX_DIMS];    /* elements */
	long    strides[BF_MAX_DIMS];  /* bytes    */
	BFbool  immutable;
	BFbool  big_endian;
	BFbool  conjugated;
} BFarray;

/* In: space, dtype, ndim, shape.  Out: data, strides.  (array.h:226-234) */
BFstatus bfArrayMalloc(BFarray* array);
BFstatus bfArrayFree(const BFarray* array);
BFstatus bfArrayCopy(const BFarray* dst, const BFarray* src);
BFstatus bfArrayMemset(const BFarray* array, int value);

/* ------------------------------------------------------------------ *
 * Device / stream glue                  (ref: src/bifrost/cuda.h:38-45)
 * All compute entry points enqueue on the calling thread's stream
 * (default: cudaStreamPerThread, ref src/cuda.cpp:34) and return without
 * waiting for the device.
 * ------------------------------------------------------------------ */
BFstatus bfStreamGet(void* stream);        /* cudaStream_t* out */
BFstatus bfStreamSet(void const* stream);  /* cudaStream_t const* in */
BFstatus bfStreamSynchronize(void);
BFstatus bfDeviceGet(int* device);
BFstatus bfDeviceSet(int device);
BFstatus bfDeviceSetById(const char* pci_bus_id);
BFstatus bfDevicesSetNoSpinCPU(void);

/* ------------------------------------------------------------------ *
 * Transpose                         (ref: src/bifrost/transpose.h:40-42)
 * out = in.transpose(axes); elements are opaque 1..16-byte words.
 * ------------------------------------------------------------------ */
BFstatus bfTranspose(BFarray const* in, BFarray const* out, int const* axes);

/* ------------------------------------------------------------------ *
 * Reduce                              (ref: src/bifrost/reduce.h:44-57)
 * Exactly one axis of `out` is shorter than in `in` by an integer factor.
 * ------------------------------------------------------------------ */
typedef enum BFreduce_op_ {
	BF_REDUCE_SUM,
	BF_REDUCE_MEAN,
	BF_REDUCE_MIN,
	BF_REDUCE_MAX,
	BF_REDUCE_STDERR,
	BF_REDUCE_POWER_SUM,
	BF_REDUCE_POWER_MEAN,
	BF_REDUCE_POWER_MIN,
	BF_REDUCE_POWER_MAX,
	BF_REDUCE_POWER_STDERR
} BFreduce_op;

BFstatus bfReduce(BFarray const* in, BFarray const* out, BFreduce_op op);

/* ------------------------------------------------------------------ *
 * FDMT                                (ref: src/bifrost/fdmt.h:48-126)
 * in  [..., nchan, ntime]  (i8/i16/i32/u8/u16/u32/f32, time fastest)
 * out [..., max_delay, ntime] f32; row d holds the transform at delay d,
 * aligned to the arrival time at the highest frequency; the last d samples
 * of row d are left untouched (ref: src/fdmt.cu:120-124,702-708).
 * Storage protocol for plan_storage / exec_storage (fdmt.h:71-78,108-115):
 *   (NULL, NULL)  library-managed;  (NULL, &size) query only;
 *   (ptr,  &size) caller-provided.
 * ------------------------------------------------------------------ */
typedef struct BFfdmt_impl* BFfdmt;

BFstatus bfFdmtCreate(BFfdmt* plan);
BFstatus bfFdmtInit(BFfdmt plan, BFsize nchan, BFsize max_delay,
                    double f0, double df, double exponent, BFspace space,
                    void* plan_storage, BFsize* plan_storage_size);
BFstatus bfFdmtSetStream(BFfdmt plan, void const* stream);
BFstatus bfFdmtExecute(BFfdmt plan, BFarray const* in, BFarray const* out,
                       BFbool negative_delays,
                       void* exec_storage, BFsize* exec_storage_size);
BFstatus bfFdmtDestroy(BFfdmt plan);

/* ------------------------------------------------------------------ *
 * FFT                                   (ref: src/bifrost/fft.h:39-60)
 * complex->complex: [i]fft; real->complex: rfft; complex->real: irfft.
 * Unnormalised in both directions.  Integer inputs are scaled to [-1,1)
 * on load (ci8/i8: x/128, ci16/i16: x/32768, ci4: nibble<<4 then /128;
 * ref: src/fft_kernels.cu:96-197).  apply_fftshift centres DC on output
 * for forward transforms and un-centres the input for inverse ones.
 * ------------------------------------------------------------------ */
typedef struct BFfft_impl* BFfft;

BFstatus bfFftCreate(BFfft* plan);
BFstatus bfFftInit(BFfft plan, BFarray const* in, BFarray const* out,
                   int rank, int const* axes, BFbool apply_fftshift,
                   size_t* tmp_storage_size);
BFstatus bfFftExecute(BFfft plan, BFarray const* in, BFarray const* out,
                      BFbool inverse, void* tmp_storage,
                      size_t tmp_storage_size);
BFstatus bfFftDestroy(BFfft plan);

/* ------------------------------------------------------------------ *
 * LinAlg                              (ref: src/bifrost/linalg.h:43-54)
 * c = alpha*a.b + beta*c;  b==NULL: a.a^H;  a==NULL: b^H.b.
 * The a.a^H / b^H.b forms write the lower triangle only (tcgen05 int8 path for
 * ci8).  a.b is a general strided SIMT product (ci8/ci16/cf32/f32 with float
 * accumulation, f64/cf64 with double; conjugated views honoured).
 * ------------------------------------------------------------------ */
typedef struct BFlinalg_impl* BFlinalg;

BFstatus bfLinAlgCreate(BFlinalg* handle);
BFstatus bfLinAlgDestroy(BFlinalg handle);
BFstatus bfLinAlgMatMul(BFlinalg handle, double alpha,
                        BFarray const* a, BFarray const* b,
                        double beta, BFarray const* c);

/* ------------------------------------------------------------------ *
 * Unpack                              (ref: src/bifrost/unpack.h:55-57)
 * 1/2/4-bit i/u/ci -> 8-bit (or wider) of the same kind.
 * ------------------------------------------------------------------ */
BFstatus bfUnpack(BFarray const* in, BFarray const* out, BFbool align_msb);

/* ------------------------------------------------------------------ *
 * Quantize                          (ref: src/bifrost/quantize.h:40-42)
 * out = IntType(rint(clip(in * scale))): f32 / cf32 in, 8/16/32-bit (complex)
 * integers out, device arrays, contiguous.
 * ------------------------------------------------------------------ */
BFstatus bfQuantize(BFarray const* in, BFarray const* out, double scale);

/* ------------------------------------------------------------------ *
 * Map                                  (ref: src/bifrost/map.h:82-94)
 * The reference JIT-compiles `func` with NVRTC.  This build ships fixed
 * sm_100a kernels for the expressions the hot-path blocks emit (detect:
 * python/bifrost/blocks/detect.py:87-136; accumulate:
 * python/bifrost/blocks/accumulate.py:67) and returns
 * BF_STATUS_UNSUPPORTED for any other expression.
 * ------------------------------------------------------------------ */
BFstatus bfMap(int ndim, long const* shape, char const* const* axis_names,
               int narg, BFarray const* const* args,
               char const* const* arg_names,
               char const* func_name, char const* func,
               char const* extra_code,
               int const* block_shape, int const* block_axes);
BFstatus bfMapClearCache(void);

/* ------------------------------------------------------------------ *
 * B200 extensions (no reference counterpart).  Same conventions: borrowed
 * BFarrays, asynchronous on the thread's stream, BFstatus returns.
 * ------------------------------------------------------------------ */

/* Stokes / power detection as a fixed kernel.  mode: 0 scalar |x|^2,
 * 1 jones (xx, yy, xy as cf32[.,2,2]...), 2 stokes (I,Q,U,V),
 * 3 stokes_i, 4 coherence.  `axis` is the polarisation axis of `in`
 * (ignored for mode 0).  Semantics of blocks/detect.py:86-138. */
BFstatus bfDetect(BFarray const* in, BFarray const* out, int mode, int axis);

/* b = beta*b + a  (blocks/accumulate.py:63-74) */
BFstatus bfAccumulate(BFarray const* a, BFarray const* b, double beta);

/* Fused GUPPI spectrometer gulp: for ci8 voltages laid out as the raw
 * GUPPI block [nchan][ntime][npol=2] (ntime = nframe*nfft), compute per
 * frame the nfft-point forward FFT with fftshift of both polarisations,
 * Stokes IQUV, sum of f_avg adjacent fine channels, and
 * out = beta*out + sum over the nframe frames.
 * out: f32 [4][nchan*nfft/f_avg].  Equivalent to the reference chain
 * transpose -> fft -> detect('stokes') -> reduce('freq', f_avg) ->
 * accumulate(nframe) (testbench/gpuspec_simple.py:44-55). */
BFstatus bfSpectrometerFused(BFarray const* in, BFarray const* out,
                             int nfft, int f_avg, double beta);

/* Host-only introspection of the FDMT plan bfFdmtInit would build (no device
 * needed).  step < 0: writes the number of steps to *nrow.  Otherwise writes
 * the row count of `step` to *nrow and, if rows != NULL, nrow triples
 * (src_row0, src_row1, delay) for step >= 1, or (row0, ndelay, 0) per channel
 * for step 0 (nchan triples). */
BFstatus bfFdmtPlanQuery(BFsize nchan, BFsize max_delay, double f0, double df,
                         double exponent, int step, int* nrow, int* rows);

/* B200 extension (test hook, host only): the work-item tables of one FDMT tile
 * pass (steps s0..s1, delay blocks of about `block_rows` rows, `nwarp` warps per
 * CTA; raw != 0: steps 1..s1 straight from a 1-byte input) as bfFdmtExecute
 * would run it -- see csrc/fdmt_tiles.cuh for the item layout.
 * header[8] = {T, nprog, nphase, slots, smem_floats, raw_bytes, edge_margin,
 * nitem} with nitem = nprog*nphase*nwarp*slots; items (if not NULL) receives
 * 4*nitem ints, aux (raw passes, if not NULL) 4*nprog*nwarp*slots ints.
 * BF_STATUS_UNSUPPORTED_SHAPE when the pass cannot be tiled. */
BFstatus bfFdmtTileQuery(BFsize nchan, BFsize max_delay, double f0, double df,
                         double exponent, int s0, int s1, int block_rows, int nwarp,
                         int raw, int* header, int* items, int* aux);

/* B200 extension (test hook, host only): the tables of pass `pass` of the
 * packed-integer FDMT schedule that bfFdmtExecute runs for 1-byte inputs --
 * see csrc/fdmt_packed.cuh for the op layout.  pass < 0: header[0] = number of
 * passes (0: the schedule does not apply to this plan).  Otherwise
 * header[24] = {s0, s1, nlev, esize, src_kind, dst_kind, T, nprog, nwarp,
 * slots, src_slots, data_bytes, lookback, nrow_out, smem_bytes, nops, lv,
 * fused, prefetch, 0...};
 * ops receives 4*nprog*nlev*nwarp*slots ints, src 4*nprog*src_slots ints,
 * hdr 4*nprog ints (each may be NULL). */
BFstatus bfFdmtPackedQuery(BFsize nchan, BFsize max_delay, double f0, double df,
                          double exponent, int pass, int* header,
                          int* ops, int* src, int* hdr);

/* B200 extension (test hook, host only): geometry of the persistent
 * single-launch form of the packed-integer schedule for a gulp of `ntime` samples.
 * header[0] = 0 when it does not apply, else header = {1, npass, lag, ipr,
 * nchunk, C, t_ref, then per pass: tb, nt, ring length of its output workspace
 * (0 for the last pass)} (7 + 3*npass longs; pass 32); tmpl (may be NULL)
 * receives ipr ints: pass << 24 | program (one item per pass, program
 * and chunk: the tiles of the program that start in the chunk). */
BFstatus bfFdmtPackedMegaQuery(BFsize nchan, BFsize max_delay, double f0, double df,
                              double exponent, long ntime, long* header, int* tmpl);

/* B200 extension (test hook): compiles the kernel bfMap would run for this
 * call (NVRTC, sm_100a) and stops; needs no device.  *mode = 0: the array
 * names are plain element references, 1: callable views (src/map.cpp:712-735
 * tries them in this order). */
BFstatus bfMapCompile(int ndim, long const* shape, char const* const* axis_names,
                      int narg, BFarray const* const* args, char const* const* arg_names,
                      char const* func_name, char const* func, char const* extra_code,
                      int const* block_axes, int* mode);

/* B200 extension: the FULL-BAND FDMT over `nrank` (2, 4 or 8) cooperating
 * plans, one per GPU (the reference runs one bfFdmt per GPU on
 * independent sub-bands only, python/bifrost/blocks/fdmt.py:59-124).  Every
 * rank calls bfFdmtInit with the full-band arguments, then bfFdmtShardInit.
 * Rank g owns the channels of sub-band g of the merge-tree step that has
 * `nrank` sub-bands (src/fdmt.cu:366-387).
 *   bfFdmtShardExecute(phase 0): the steps up to the split step on the rank's
 *     own channels (`in` = [nchan/nrank][ntime] i8/u8), into the rank's block
 *     of rows of the split-step workspace;
 *   the caller exchanges the blocks (NCCL all-gather over NVLink; layout from
 *     bfFdmtShardQuery: same byte offsets on every rank);
 *   bfFdmtShardExecute(phase 1): the rank's delay blocks of the remaining
 *     steps into `out` = [max_delay][ntime] f32 (other rows untouched).
 * All ranks' phase-1 rows together equal bfFdmtExecute's output bit for bit.
 * Workspace protocol as bfFdmtExecute (exec_storage NULL: size query). */
BFstatus bfFdmtShardInit(BFfdmt plan, int rank, int nrank);
BFstatus bfFdmtShardExecute(BFfdmt plan, int phase, BFarray const* in, BFarray const* out,
                            void* exec_storage, BFsize* exec_storage_size);
/* Phase 1 without an exchange: the split-step rows are staged by TMA straight
 * from the workspace of the rank that produced them.  peer_storage[g] = rank g's
 * exec_storage as mapped into this process (CUDA IPC / NVLink peer access;
 * the own rank's entry is exec_storage).  The caller orders it after every
 * rank's phase 0 and keeps the workspaces untouched until every rank is done. */
BFstatus bfFdmtShardExecutePeers(BFfdmt plan, BFarray const* in, BFarray const* out,
                                 void* exec_storage, BFsize* exec_storage_size,
                                 void const* const* peer_storage, int npeer);
/* info: [0] byte offset of the split-step rows in the workspace, [1] row pitch
 * (bytes), [2] rows, [3] bytes per element, [4] nrank, [5] split step,
 * [6 .. 6+nrank] first row of each rank's block and the end; then n and n
 * triples (first delay, delays, owning rank) of the last pass's delay blocks.
 * *ninfo: capacity (longs) in, used out. */
BFstatus bfFdmtShardQuery(BFfdmt plan, long ntime, long* info, int* ninfo);

/* B200 extension: CUDA IPC for bfMalloc'ed device buffers (one process per GPU
 * mapping each other's buffers; used by the sharded FDMT's peer-access phase).
 * handle64: the 64 bytes of a cudaIpcMemHandle_t; ptr: start of the allocation. */
BFstatus bfIpcGetHandle(void* ptr, void* handle64);
BFstatus bfIpcOpenHandle(void const* handle64, void** ptr);
BFstatus bfIpcCloseHandle(void* ptr);

/* ------------------------------------------------------------------ *
 * Rings                               (ref: src/bifrost/ring.h:60-227)
 * Byte-addressed circular buffers between blocks, in any memory space, with
 * a ghost region that makes every span contiguous (device-to-device ghost
 * copies on the caller's stream for rings in CUDA space).  One writer stream
 * of sequences (name, time tag, header) and spans; readers open sequences by
 * name / time tag / earliest / latest, optionally *guaranteed* (the writer
 * blocks instead of overwriting what they still hold).  Blocking calls
 * return BF_STATUS_END_OF_DATA when the data they wait for can no longer
 * come, non-blocking reservations BF_STATUS_WOULD_BLOCK.
 * BFrsequence / BFwsequence handles are valid BFsequence handles, BFrspan /
 * BFwspan handles valid BFspan handles.
 * ------------------------------------------------------------------ */
typedef struct BFring_impl*        BFring;
typedef struct BFsequence_wrapper* BFsequence;
typedef struct BFrsequence_impl*   BFrsequence;
typedef struct BFwsequence_impl*   BFwsequence;
typedef struct BFspan_impl*        BFspan;
typedef struct BFrspan_impl*       BFrspan;
typedef struct BFwspan_impl*       BFwspan;

BFstatus bfRingCreate(BFring* ring, const char* name, BFspace space);
BFstatus bfRingDestroy(BFring ring);
/* grow-only: at least `contiguous_bytes` per span, `capacity_bytes` buffered
 * (rounded up to a power of two), `nringlet` ringlets; contents are kept */
BFstatus bfRingResize(BFring ring, BFsize contiguous_bytes, BFsize capacity_bytes, BFsize nringlet);
BFstatus bfRingGetName(BFring ring, const char** name);
BFstatus bfRingGetSpace(BFring ring, BFspace* space);
/* later (re)allocations of a host-space ring come from the NUMA node of
 * `core` (-1: no preference).  The reference needs hwloc for this. */
BFstatus bfRingSetAffinity(BFring ring, int core);
BFstatus bfRingGetAffinity(BFring ring, int* core);
BFstatus bfRingLock(BFring ring);
BFstatus bfRingUnlock(BFring ring);
BFstatus bfRingLockedGetData(BFring ring, void** data);
BFstatus bfRingLockedGetContiguousSpan(BFring ring, BFsize* val);
BFstatus bfRingLockedGetTotalSpan(BFring ring, BFsize* val);
BFstatus bfRingLockedGetNRinglet(BFring ring, BFsize* val);
BFstatus bfRingLockedGetStride(BFring ring, BFsize* val);
BFstatus bfRingBeginWriting(BFring ring);
BFstatus bfRingEndWriting(BFring ring);
BFstatus bfRingWritingEnded(BFring ring, BFbool* writing_ended);
/* name: unique among the sequences in the ring, or ""; time_tag: unique, or
 * BFoffset(-1) */
BFstatus bfRingSequenceBegin(BFwsequence* sequence, BFring ring, const char* name,
                             BFoffset time_tag, BFsize header_size, const void* header,
                             BFsize nringlet, BFoffset offset_from_head);
BFstatus bfRingSequenceEnd(BFwsequence sequence, BFoffset offset_from_head);
BFstatus bfRingSequenceOpen(BFrsequence* sequence, BFring ring, const char* name, BFbool guarantee);
BFstatus bfRingSequenceOpenAt(BFrsequence* sequence, BFring ring, BFoffset time_tag, BFbool guarantee);
BFstatus bfRingSequenceOpenLatest(BFrsequence* sequence, BFring ring, BFbool guarantee);
BFstatus bfRingSequenceOpenEarliest(BFrsequence* sequence, BFring ring, BFbool guarantee);
BFstatus bfRingSequenceNext(BFrsequence sequence);
BFstatus bfRingSequenceClose(BFrsequence sequence);
BFstatus bfRingSequenceGetRing(BFsequence sequence, BFring* ring);
BFstatus bfRingSequenceGetName(BFsequence sequence, const char** name);
BFstatus bfRingSequenceGetTimeTag(BFsequence sequence, BFoffset* time_tag);
BFstatus bfRingSequenceGetHeader(BFsequence sequence, const void** hdr);
BFstatus bfRingSequenceGetHeaderSize(BFsequence sequence, BFsize* size);
BFstatus bfRingSequenceGetNRinglet(BFsequence sequence, BFsize* nringlet);
typedef struct BFsequence_info_ {
	BFring      ring;
	const char* name;
	BFoffset    time_tag;
	const void* header;
	BFsize      header_size;
	BFsize      nringlet;
} BFsequence_info;
BFstatus bfRingSequenceGetInfo(BFsequence sequence, BFsequence_info* sequence_info);
BFstatus bfRingSpanReserve(BFwspan* span, BFring ring, BFsize size, BFbool nonblocking);
BFstatus bfRingSpanCommit(BFwspan span, BFsize size);
BFstatus bfRingSpanAcquire(BFrspan* span, BFrsequence sequence, BFoffset offset, BFsize size);
BFstatus bfRingSpanRelease(BFrspan span);
BFstatus bfRingSpanGetSizeOverwritten(BFrspan span, BFsize* val);
BFstatus bfRingSpanGetRing(BFspan span, BFring* ring);
BFstatus bfRingSpanGetData(BFspan span, void** data);
BFstatus bfRingSpanGetSize(BFspan span, BFsize* val);
BFstatus bfRingSpanGetStride(BFspan span, BFsize* val);
BFstatus bfRingSpanGetOffset(BFspan span, BFsize* val);
BFstatus bfRingSpanGetNRinglet(BFspan span, BFsize* val);
typedef struct BFspan_info_ {
	BFring      ring;
	void*       data;
	BFsize      size;
	BFsize      stride;
	BFsize      offset;
	BFsize      nringlet;
} BFspan_info;
BFstatus bfRingSpanGetInfo(BFspan span, BFspan_info* span_info);

/* ------------------------------------------------------------------ *
 * Process log                        (ref: src/bifrost/proclog.h:41-48)
 * One text file per log under BF_PROCLOG_DIR/<pid>/<name> ("block/quantity"),
 * rewritten on every update, removed with the handle / the process.
 * The directory can be moved with $BIFROST_B200_PROCLOG_DIR.
 * ------------------------------------------------------------------ */
#define BFB_PROCLOG_DIR "/dev/shm/bifrost"
typedef struct BFproclog_impl* BFproclog;
BFstatus bfProcLogCreate(BFproclog* log_ptr, const char* name);
BFstatus bfProcLogDestroy(BFproclog log);
BFstatus bfProcLogUpdate(BFproclog log, const char* str);

/* ------------------------------------------------------------------ *
 * Thread placement                  (ref: src/bifrost/affinity.h:41-46)
 * core = -1 unbinds; GetCore gives -1 for a thread that may run on several.
 * ------------------------------------------------------------------ */
BFstatus bfAffinitySetCore(int core);
BFstatus bfAffinityGetCore(int* core);
BFstatus bfAffinitySetOpenMPCores(BFsize nthread, const int* thread_cores);

/* Number of kernels this library has launched since load (all threads). */
BFstatus bfGetLaunchCount(unsigned long long* count);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* BIFROST_B200_H_ */
