// quantize.cu -- bfQuantize for device arrays (replaces the element loop of
// src/quantize.cpp:45-90, which is CPU-only in the reference; its GPU twin
// there lives in guantize.cu).  out = IntType(rint(clip(in * scale))) with the
// clip range [-max, +max] for signed outputs (the minimum two's-complement
// value is never produced) and [0, max] for unsigned ones; rint rounds half to
// even.  f32 / cf32 in (a complex sample is two floats), 8/16/32-bit integer
// out; contiguous arrays.
// Sub-byte outputs (guantize.cu:146-348): 4-bit values clip to [-7, 7] and two
// of them share a byte, the FIRST in the high nibble; 2-bit values clip to
// [-1, 1], four per byte, first in bits 7:6; 1-bit values are (x*scale >= 0),
// eight per byte, first in bit 7.  (The reference's 1-bit masks for its first
// three samples -- 0x08, 0x04, 0x02 applied to bits 7, 6, 5 -- drop them; the
// other five land where they do here.  DESIGN.md section 8.)
#include "core.hpp"

namespace bfb {

template<typename O> struct QRange;
template<> struct QRange<signed char>    { static constexpr float lo = -127.f,        hi = 127.f; };
template<> struct QRange<short>          { static constexpr float lo = -32767.f,      hi = 32767.f; };
template<> struct QRange<int>            { static constexpr float lo = -2147483647.f, hi = 2147483647.f; };
template<> struct QRange<unsigned char>  { static constexpr float lo = 0.f, hi = 255.f; };
template<> struct QRange<unsigned short> { static constexpr float lo = 0.f, hi = 65535.f; };
template<> struct QRange<unsigned int>   { static constexpr float lo = 0.f, hi = 4294967295.f; };

template<typename O>
__device__ __forceinline__ O quantize_one(float x, float scale) {
	float v = __fmul_rn(x, scale);
	v = fminf(fmaxf(v, QRange<O>::lo), QRange<O>::hi);
	// 32-bit limits are not exactly representable: convert through double
	return (O)rint((double)v);
}

template<typename O>
__global__ void __launch_bounds__(256)
quantize_kernel(const float* __restrict__ in, O* __restrict__ out, long n, float scale, bool vec) {
	long gstride = (long)gridDim.x * blockDim.x;
	if( vec ) {
		// four values per thread: one 16-byte load, one 4/8/16-byte store
		for( long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += gstride ) {
			const float4 v = *(const float4*)(in + 4 * i);
			struct __align__(4 * sizeof(O)) Vec { O a, b, c, d; };
			Vec o = {quantize_one<O>(v.x, scale), quantize_one<O>(v.y, scale),
			         quantize_one<O>(v.z, scale), quantize_one<O>(v.w, scale)};
			*(Vec*)(out + 4 * i) = o;
		}
		for( long i = (n & ~3L) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gstride )
			out[i] = quantize_one<O>(in[i], scale);
	} else {
		for( long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gstride )
			out[i] = quantize_one<O>(in[i], scale);
	}
}

// nbit = 4, 2 or 1: one output byte per thread-iteration from 8/nbit floats
// (two aligned float4 loads where the input allows it).
template<int NBIT>
__device__ __forceinline__ unsigned quantize_sub(float x, float scale) {
	float v = __fmul_rn(x, scale);
	if( NBIT == 1 ) return v >= 0.f ? 1u : 0u;
	const float lim = NBIT == 4 ? 7.f : 1.f;
	v = fminf(fmaxf(v, -lim), lim);
	return (unsigned)(int)rintf(v) & ((1u << NBIT) - 1u);
}
template<int NBIT>
__global__ void __launch_bounds__(256)
quantize_sub_kernel(const float* __restrict__ in, unsigned char* __restrict__ out, long nbyte, float scale, bool vec) {
	constexpr int PER = 8 / NBIT;
	long gstride = (long)gridDim.x * blockDim.x;
	for( long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nbyte; i += gstride ) {
		float x[PER];
		if( vec && PER >= 4 ) {
#pragma unroll
			for( int q=0; q<PER/4; ++q ) {
				const float4 v = *(const float4*)(in + PER * i + 4 * q);
				x[4*q] = v.x; x[4*q+1] = v.y; x[4*q+2] = v.z; x[4*q+3] = v.w;
			}
		} else if( vec ) {
			const float2 v = *(const float2*)(in + PER * i);
			x[0] = v.x; x[1] = v.y;
		} else {
#pragma unroll
			for( int k=0; k<PER; ++k ) x[k] = in[PER * i + k];
		}
		unsigned b = 0;
#pragma unroll
		for( int k=0; k<PER; ++k ) b |= quantize_sub<NBIT>(x[k], scale) << (8 - NBIT * (k + 1));
		out[i] = (unsigned char)b;
	}
}
template<int NBIT>
static BFstatus launch_quantize_sub(const void* in, void* out, long n, double scale) {
	const long nbyte = n * NBIT / 8;
	bool vec = (uintptr_t)in % 16 == 0;
	unsigned grid = (unsigned)std::min<long>(div_up<long>(std::max<long>(nbyte, 1), 256), 148L * 16);
	quantize_sub_kernel<NBIT><<<grid, 256, 0, thread_stream()>>>((const float*)in, (unsigned char*)out, nbyte, (float)scale, vec);
	count_launch();
	BFB_CUDA(cudaGetLastError(), BF_STATUS_INTERNAL_ERROR);
	return BF_STATUS_SUCCESS;
}

template<typename O>
static BFstatus launch_quantize(const void* in, void* out, long n, double scale) {
	bool vec = (uintptr_t)in % 16 == 0 && (uintptr_t)out % (4 * sizeof(O)) == 0;
	unsigned grid = (unsigned)std::min<long>(div_up<long>(std::max<long>(n / 4, 1), 256), 148L * 16);
	quantize_kernel<O><<<grid, 256, 0, thread_stream()>>>((const float*)in, (O*)out, n, (float)scale, vec);
	count_launch();
	BFB_CUDA(cudaGetLastError(), BF_STATUS_INTERNAL_ERROR);
	return BF_STATUS_SUCCESS;
}

} // namespace bfb

using namespace bfb;

extern "C"
BFstatus bfQuantize(BFarray const* in, BFarray const* out, double scale) {
	BFB_ASSERT(in && out, BF_STATUS_INVALID_POINTER);
	BFB_ASSERT(!out->immutable, BF_STATUS_INVALID_POINTER);
	BFB_ASSERT(space_on_device(in->space) && space_on_device(out->space), BF_STATUS_UNSUPPORTED_SPACE);
	BFB_ASSERT(in->ndim == out->ndim && in->ndim >= 1 && in->ndim <= BF_MAX_DIMS, BF_STATUS_INVALID_SHAPE);
	for( int d=0; d<in->ndim; ++d ) BFB_ASSERT(in->shape[d] == out->shape[d], BF_STATUS_INVALID_SHAPE);
	BFB_ASSERT(in->dtype == BF_DTYPE_F32 || in->dtype == BF_DTYPE_CF32, BF_STATUS_UNSUPPORTED_DTYPE);
	BFB_ASSERT(dtype_is_complex(in->dtype) == dtype_is_complex(out->dtype), BF_STATUS_INVALID_DTYPE);
	BFB_ASSERT(array_is_contiguous(in), BF_STATUS_UNSUPPORTED_STRIDE);
	if( dtype_nbit(out->dtype) >= 8 ) BFB_ASSERT(array_is_contiguous(out), BF_STATUS_UNSUPPORTED_STRIDE);
	else {
		// packed output: contiguity in bits (the last dim's stride may be 0 or 1 byte)
		long expect_bits = dtype_nbit(out->dtype);
		for( int d=out->ndim-1; d>=0; --d ) {
			if( out->shape[d] != 1 && d != out->ndim-1 )
				BFB_ASSERT(out->strides[d] * 8 == expect_bits, BF_STATUS_UNSUPPORTED_STRIDE);
			expect_bits *= out->shape[d];
		}
	}
	long n = dtype_is_complex(in->dtype) ? 2 : 1;
	for( int d=0; d<in->ndim; ++d ) n *= in->shape[d];
	if( n == 0 ) return BF_STATUS_SUCCESS;
	switch( out->dtype ) {
	// (whole bytes only, as quantize.cpp:305,329,353 asserts)
	case BF_DTYPE_I4: case BF_DTYPE_CI4: BFB_ASSERT(n % 2 == 0, BF_STATUS_INVALID_SHAPE); return launch_quantize_sub<4>(in->data, out->data, n, scale);
	case BF_DTYPE_I2: case BF_DTYPE_CI2: BFB_ASSERT(n % 4 == 0, BF_STATUS_INVALID_SHAPE); return launch_quantize_sub<2>(in->data, out->data, n, scale);
	case BF_DTYPE_I1: case BF_DTYPE_CI1: BFB_ASSERT(n % 8 == 0, BF_STATUS_INVALID_SHAPE); return launch_quantize_sub<1>(in->data, out->data, n, scale);
	case BF_DTYPE_I8:  case BF_DTYPE_CI8:  return launch_quantize<signed char>(in->data, out->data, n, scale);
	case BF_DTYPE_I16: case BF_DTYPE_CI16: return launch_quantize<short>(in->data, out->data, n, scale);
	case BF_DTYPE_I32: case BF_DTYPE_CI32: return launch_quantize<int>(in->data, out->data, n, scale);
	case BF_DTYPE_U8:  return launch_quantize<unsigned char>(in->data, out->data, n, scale);
	case BF_DTYPE_U16: return launch_quantize<unsigned short>(in->data, out->data, n, scale);
	case BF_DTYPE_U32: return launch_quantize<unsigned int>(in->data, out->data, n, scale);
	default: BFB_FAIL(BF_STATUS_UNSUPPORTED_DTYPE);
	}
}
