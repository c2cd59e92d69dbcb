// fdmt.cu -- bfFdmt* for sm_100a.
//
// Replaces: src/fdmt.cu:724-814 (C entry points), :531-628 (storage
// protocol), :629-718 (execute), :52-92 (init kernel), :95-155 (step kernel).
// The host plan lives in fdmt_plan.hpp.
//
// Arithmetic contract (bit-exact with the reference, which is itself a fixed
// tree of fp32 operations):
//   state0[row0(c)+d][t] = (sum_{k<=d} float(in[c][t-k])) * (1.f/(d+1))   t >= d
//                        = NaN                                            t <  d
//   state_s[r][t]        = state_{s-1}[src0][t] + (t >= delay ?
//                          state_{s-1}[src1][t-delay] : 0)       (absent src = 0)
//   out[r][t-r]          = state_last[r][t]                      t >= r
// (negative_delays mirrors the time axis as the reference does; where the
// reference indexes before the start of a row -- src/fdmt.cu:80-81 with
// reverse_time -- this implementation contributes 0 instead of reading out
// of bounds.)
//
// All index arithmetic is 64-bit (the reference's 32-bit `t + ostride*row`
// overflows at 4096 chan x 128k samples).
#include "core.hpp"
#include "shape.hpp"
#include "fdmt_plan.hpp"

#include <math_constants.h>
#include <algorithm>
#include <vector>

namespace bfb {

enum { FDMT_TIME_ALIGN = 128 };   // state row pitch is a multiple of this

template<typename T> struct Vec4 { T v[4]; };

// ---------------------------------------------------------------------------
// Step 0.  One thread produces 4 consecutive time samples of every delay row
// of one channel.
template<typename In>
__global__ void __launch_bounds__(256)
fdmt_init_kernel(const In* __restrict__ in, long istride, long ibatchstride,
                 float* __restrict__ state, long sstride, long sbatchstride,
                 const int* __restrict__ row_offsets,
                 int nchan, long ntime, bool reverse_band, bool reverse_time) {
	int  c = blockIdx.y;
	int  b = blockIdx.z;
	long t0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
	if( t0 >= ntime ) return;
	int row0   = row_offsets[c];
	int ndelay = row_offsets[c+1] - row0;
	int c_in = reverse_band ? nchan-1 - c : c;
	const In* src = in + (long)c_in * istride + (long)b * ibatchstride;
	float* dst = state + (long)row0 * sstride + (long)b * sbatchstride;
	float acc[4] = {0.f, 0.f, 0.f, 0.f};
	for( int d=0; d<ndelay; ++d ) {
		float scale = 1.f / (d + 1);
		float o[4];
#pragma unroll
		for( int j=0; j<4; ++j ) {
			long t = t0 + j;
			float val = CUDART_NAN_F;
			if( t < ntime && t >= d ) {
				long ti = (reverse_time ? ntime-1 - t : t) - d;
				acc[j] += (ti >= 0) ? (float)src[ti] : 0.f;
				val = acc[j] * scale;
			}
			o[j] = val;
		}
		// sstride is a multiple of 4 and t0 % 4 == 0: aligned 128-bit store
		// (the pad beyond ntime is scratch).
		*(float4*)(dst + (long)d * sstride + t0) = make_float4(o[0], o[1], o[2], o[3]);
	}
}

// ---------------------------------------------------------------------------
// Merge step.  One thread produces 4 consecutive time samples of one row.
template<bool FINAL>
__global__ void __launch_bounds__(256)
fdmt_step_kernel(const float* __restrict__ prev, long pstride, long pbatchstride,
                 float* __restrict__ next, long nstride, long nbatchstride,
                 const int4* __restrict__ rows,   // (src0, src1, delay, -)
                 long ntime, bool reverse_time) {
	int  r = blockIdx.y;
	int  b = blockIdx.z;
	long t0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
	if( t0 >= ntime ) return;
	int4 row = rows[r];
	const float* a  = prev + (long)b * pbatchstride + (long)row.x * pstride;
	const float* bb = prev + (long)b * pbatchstride + (long)row.y * pstride;
	int delay = row.z;
	float4 va = (row.x >= 0) ? *(const float4*)(a + t0) : make_float4(0.f, 0.f, 0.f, 0.f);
	float o[4] = {va.x, va.y, va.z, va.w};
	if( row.y >= 0 ) {
#pragma unroll
		for( int j=0; j<4; ++j ) {
			long t = t0 + j;
			if( t >= delay && t < ntime ) o[j] += bb[t - delay];
		}
	}
	float* dst = next + (long)b * nbatchstride;
	if( !FINAL ) {
		*(float4*)(dst + (long)r * nstride + t0) = make_float4(o[0], o[1], o[2], o[3]);
	} else {
		// Diagonal re-indexing: output row r is shifted so that column t'
		// holds arrival time t'+r (ref: src/fdmt.cu:702-708,150-151).
#pragma unroll
		for( int j=0; j<4; ++j ) {
			long t = t0 + j;
			if( t < ntime && t >= r ) {
				long to = reverse_time ? (ntime-1 - t) + r : t - r;
				dst[(long)r * nstride + to] = o[j];
			}
		}
	}
}

} // namespace bfb

using namespace bfb;

struct BFfdmt_impl {
	FdmtPlan plan;
	bool     planned = false;
	cudaStream_t stream = nullptr;
	bool     stream_set = false;
	// device plan
	void*  own_plan_storage = nullptr;
	size_t own_plan_size = 0;
	int*   d_row_offsets = nullptr;      // [nchan+1]
	int4*  d_rows = nullptr;             // [nstep][plan_stride]
	long   plan_stride = 0;
	// exec workspace
	void*  own_exec_storage = nullptr;
	size_t own_exec_size = 0;

	cudaStream_t get_stream() { return stream_set ? stream : thread_stream(); }

	size_t plan_bytes() const {
		size_t off = round_up<size_t>((plan.nchan + 1) * sizeof(int), 512);
		off += (size_t)plan.nstep() * plan_stride * sizeof(int4);
		return off;
	}
	~BFfdmt_impl() {
		if( own_plan_storage ) cudaFree(own_plan_storage);
		if( own_exec_storage ) cudaFree(own_exec_storage);
	}
};

extern "C" {

BFstatus bfFdmtCreate(BFfdmt* plan_ptr) {
	BFB_ASSERT(plan_ptr, BF_STATUS_INVALID_POINTER);
	*plan_ptr = nullptr;
	BFB_TRY(*plan_ptr = new BFfdmt_impl());
	return BF_STATUS_SUCCESS;
}

BFstatus bfFdmtDestroy(BFfdmt plan) {
	BFB_ASSERT(plan, BF_STATUS_INVALID_HANDLE);
	delete plan;
	return BF_STATUS_SUCCESS;
}

BFstatus bfFdmtSetStream(BFfdmt plan, void const* stream) {
	BFB_ASSERT(plan,   BF_STATUS_INVALID_HANDLE);
	BFB_ASSERT(stream, BF_STATUS_INVALID_POINTER);
	plan->stream = *(cudaStream_t const*)stream;
	plan->stream_set = true;
	return BF_STATUS_SUCCESS;
}

BFstatus bfFdmtInit(BFfdmt plan, BFsize nchan, BFsize max_delay,
                    double f0, double df, double exponent, BFspace space,
                    void* plan_storage, BFsize* plan_storage_size) {
	BFB_ASSERT(plan, BF_STATUS_INVALID_HANDLE);
	BFB_ASSERT(nchan > 1, BF_STATUS_INVALID_ARGUMENT);
	BFB_ASSERT(max_delay >= 1, BF_STATUS_INVALID_ARGUMENT);
	BFB_ASSERT(space_on_device(space), BF_STATUS_UNSUPPORTED_SPACE);
	bool ok = false;
	BFB_TRY(ok = plan->plan.build((int)nchan, (int)max_delay, f0, df, exponent));
	BFB_ASSERT(ok, BF_STATUS_INTERNAL_ERROR);
	plan->planned = true;
	plan->plan_stride = round_up<long>(plan->plan.nrow_max, 128);
	size_t need = plan->plan_bytes();
	if( plan_storage_size ) {
		if( !plan_storage ) { *plan_storage_size = need; return BF_STATUS_SUCCESS; }
		BFB_ASSERT(*plan_storage_size >= need, BF_STATUS_INSUFFICIENT_STORAGE);
	} else {
		BFB_ASSERT(!plan_storage, BF_STATUS_INVALID_ARGUMENT);
		if( plan->own_plan_size < need ) {
			if( plan->own_plan_storage ) cudaFree(plan->own_plan_storage);
			plan->own_plan_storage = nullptr; plan->own_plan_size = 0;
			BFB_CUDA(cudaMalloc(&plan->own_plan_storage, need), BF_STATUS_MEM_ALLOC_FAILED);
			plan->own_plan_size = need;
		}
		plan_storage = plan->own_plan_storage;
	}
	char* base = (char*)plan_storage;
	plan->d_row_offsets = (int*)base;
	plan->d_rows = (int4*)(base + round_up<size_t>((nchan + 1) * sizeof(int), 512));
	// Upload: row offsets of step 0, then one padded int4 table per step.
	FdmtPlan const& P = plan->plan;
	std::vector<int> offsets(nchan + 1);
	for( size_t c=0; c<nchan; ++c ) offsets[c] = P.bands[0][c].row0;
	offsets[nchan] = P.nrow(0);
	std::vector<int4> table((size_t)P.nstep() * plan->plan_stride, make_int4(-1, -1, 0, 0));
	for( int s=1; s<P.nstep(); ++s ) {
		for( size_t r=0; r<P.rows[s].size(); ++r ) {
			FdmtRow const& row = P.rows[s][r];
			table[(size_t)s * plan->plan_stride + r] = make_int4(row.src0, row.src1, row.delay, 0);
		}
	}
	cudaStream_t st = plan->get_stream();
	BFB_CUDA(cudaMemcpyAsync(plan->d_row_offsets, offsets.data(), offsets.size()*sizeof(int),
	                         cudaMemcpyHostToDevice, st), BF_STATUS_MEM_OP_FAILED);
	BFB_CUDA(cudaMemcpyAsync(plan->d_rows, table.data(), table.size()*sizeof(int4),
	                         cudaMemcpyHostToDevice, st), BF_STATUS_MEM_OP_FAILED);
	BFB_CUDA(cudaStreamSynchronize(st), BF_STATUS_DEVICE_ERROR);
	return BF_STATUS_SUCCESS;
}

BFstatus bfFdmtPlanQuery(BFsize nchan, BFsize max_delay, double f0, double df,
                         double exponent, int step, int* nrow, int* rows) {
	BFB_ASSERT(nrow, BF_STATUS_INVALID_POINTER);
	BFB_ASSERT(nchan > 1 && max_delay >= 1, BF_STATUS_INVALID_ARGUMENT);
	FdmtPlan P;
	bool ok = false;
	BFB_TRY(ok = P.build((int)nchan, (int)max_delay, f0, df, exponent));
	BFB_ASSERT(ok, BF_STATUS_INTERNAL_ERROR);
	if( step < 0 ) { *nrow = P.nstep(); return BF_STATUS_SUCCESS; }
	BFB_ASSERT(step < P.nstep(), BF_STATUS_INVALID_ARGUMENT);
	*nrow = P.nrow(step);
	if( rows ) {
		if( step == 0 ) {
			for( size_t c=0; c<P.bands[0].size(); ++c ) {
				rows[3*c+0] = P.bands[0][c].row0;
				rows[3*c+1] = P.bands[0][c].ndelay;
				rows[3*c+2] = 0;
			}
		} else {
			for( size_t r=0; r<P.rows[step].size(); ++r ) {
				rows[3*r+0] = P.rows[step][r].src0;
				rows[3*r+1] = P.rows[step][r].src1;
				rows[3*r+2] = P.rows[step][r].delay;
			}
		}
	}
	return BF_STATUS_SUCCESS;
}

BFstatus bfFdmtExecute(BFfdmt plan, BFarray const* in, BFarray const* out,
                       BFbool negative_delays,
                       void* exec_storage, BFsize* exec_storage_size) {
	BFB_ASSERT(plan, BF_STATUS_INVALID_HANDLE);
	BFB_ASSERT(in,   BF_STATUS_INVALID_POINTER);
	BFB_ASSERT(out,  BF_STATUS_INVALID_POINTER);
	BFB_ASSERT(plan->planned, BF_STATUS_INVALID_STATE);
	FdmtPlan const& P = plan->plan;
	int ndim = in->ndim;
	BFB_ASSERT(ndim == out->ndim && ndim >= 2 && ndim <= BF_MAX_DIMS, BF_STATUS_INVALID_SHAPE);
	BFB_ASSERT( in->shape[ndim-2] == P.nchan,     BF_STATUS_INVALID_SHAPE);
	BFB_ASSERT(out->shape[ndim-2] == P.max_delay, BF_STATUS_INVALID_SHAPE);
	BFB_ASSERT( in->shape[ndim-1] == out->shape[ndim-1], BF_STATUS_INVALID_SHAPE);
	long ntime = in->shape[ndim-1];
	// Batch dims must fuse to a single strided batch dim on both sides
	// (ref: src/fdmt.cu:779-796).
	long nbatch = 1, ibatchbytes = 0, obatchbytes = 0;
	if( ndim > 2 ) {
		StridedView v[2];
		v[0].ndim = v[1].ndim = ndim - 2;
		for( int d=0; d<ndim-2; ++d ) {
			BFB_ASSERT(in->shape[d] == out->shape[d], BF_STATUS_INVALID_SHAPE);
			v[0].shape[d] = v[1].shape[d] = in->shape[d];
			v[0].strides[d] = in->strides[d];
			v[1].strides[d] = out->strides[d];
		}
		merge_views(v, 2);
		BFB_ASSERT(v[0].ndim == 1, BF_STATUS_UNSUPPORTED_SHAPE);
		nbatch = v[0].shape[0];
		ibatchbytes = v[0].strides[0];
		obatchbytes = v[1].strides[0];
	}
	long sstride = round_up<long>(ntime, FDMT_TIME_ALIGN);
	long sbatchstride = (long)P.nrow_max * sstride;
	size_t need = 2 * (size_t)nbatch * sbatchstride * sizeof(float);
	if( exec_storage_size ) {
		if( !exec_storage ) { *exec_storage_size = need; return BF_STATUS_SUCCESS; }
		BFB_ASSERT(*exec_storage_size >= need, BF_STATUS_INSUFFICIENT_STORAGE);
	} else {
		BFB_ASSERT(!exec_storage, BF_STATUS_INVALID_ARGUMENT);
		if( plan->own_exec_size < need ) {
			if( plan->own_exec_storage ) cudaFree(plan->own_exec_storage);
			plan->own_exec_storage = nullptr; plan->own_exec_size = 0;
			BFB_CUDA(cudaMalloc(&plan->own_exec_storage, need), BF_STATUS_MEM_ALLOC_FAILED);
			plan->own_exec_size = need;
		}
		exec_storage = plan->own_exec_storage;
	}
	BFB_ASSERT(space_on_device(in->space),  BF_STATUS_UNSUPPORTED_SPACE);
	BFB_ASSERT(space_on_device(out->space), BF_STATUS_UNSUPPORTED_SPACE);
	BFB_ASSERT(out->dtype == BF_DTYPE_F32,  BF_STATUS_UNSUPPORTED_DTYPE);
	long isize = dtype_nbyte(in->dtype);
	BFB_ASSERT(isize > 0, BF_STATUS_UNSUPPORTED_DTYPE);
	BFB_ASSERT( in->strides[ndim-1] == isize, BF_STATUS_UNSUPPORTED_STRIDE);
	BFB_ASSERT(out->strides[ndim-1] == 4,     BF_STATUS_UNSUPPORTED_STRIDE);
	BFB_ASSERT( in->strides[ndim-2] > 0 &&  in->strides[ndim-2] % isize == 0, BF_STATUS_UNSUPPORTED_STRIDE);
	BFB_ASSERT(out->strides[ndim-2] > 0 && out->strides[ndim-2] % 4 == 0,     BF_STATUS_UNSUPPORTED_STRIDE);
	BFB_ASSERT(ibatchbytes % isize == 0 && obatchbytes % 4 == 0,              BF_STATUS_UNSUPPORTED_STRIDE);
	if( ntime == 0 || nbatch == 0 ) return BF_STATUS_SUCCESS;
	BFB_ASSERT(nbatch <= 65535 && P.nrow_max <= 65535, BF_STATUS_UNSUPPORTED_SHAPE);
	long istride = in->strides[ndim-2] / isize,  ibatch = ibatchbytes / isize;
	long ostride = out->strides[ndim-2] / 4,     obatch = obatchbytes / 4;

	float* buf_a = (float*)exec_storage;
	float* buf_b = buf_a + (size_t)nbatch * sbatchstride;
	bool rev = negative_delays != 0;
	cudaStream_t st = plan->get_stream();
	dim3 block(256);
	unsigned gx = (unsigned)div_up<long>(div_up<long>(ntime, 4), 256);
	dim3 grid0(gx, P.nchan, (unsigned)nbatch);
#define BFB_FDMT_INIT(T) \
	fdmt_init_kernel<T><<<grid0, block, 0, st>>>((const T*)in->data, istride, ibatch, \
		buf_a, sstride, sbatchstride, plan->d_row_offsets, P.nchan, ntime, \
		P.reverse_band, rev)
	switch( in->dtype ) {
	case BF_DTYPE_I8:  BFB_FDMT_INIT(int8_t);   break;
	case BF_DTYPE_I16: BFB_FDMT_INIT(int16_t);  break;
	case BF_DTYPE_I32: BFB_FDMT_INIT(int32_t);  break;
	case BF_DTYPE_U8:  BFB_FDMT_INIT(uint8_t);  break;
	case BF_DTYPE_U16: BFB_FDMT_INIT(uint16_t); break;
	case BF_DTYPE_U32: BFB_FDMT_INIT(uint32_t); break;
	case BF_DTYPE_F32: BFB_FDMT_INIT(float);    break;
	default: BFB_FAIL(BF_STATUS_UNSUPPORTED_DTYPE);
	}
#undef BFB_FDMT_INIT
	count_launch();
	float* cur = buf_a;
	float* nxt = buf_b;
	int nstep = P.nstep();
	for( int s=1; s<nstep; ++s ) {
		dim3 grid(gx, P.nrow(s), (unsigned)nbatch);
		const int4* rows = plan->d_rows + (size_t)s * plan->plan_stride;
		if( s == nstep-1 ) {
			fdmt_step_kernel<true><<<grid, block, 0, st>>>(cur, sstride, sbatchstride,
				(float*)out->data, ostride, obatch, rows, ntime, rev);
		} else {
			fdmt_step_kernel<false><<<grid, block, 0, st>>>(cur, sstride, sbatchstride,
				nxt, sstride, sbatchstride, rows, ntime, rev);
		}
		count_launch();
		std::swap(cur, nxt);
	}
	BFB_CUDA(cudaGetLastError(), BF_STATUS_INTERNAL_ERROR);
	return BF_STATUS_SUCCESS;
}

} // extern "C"
