// spectrometer.cu -- bfSpectrometerFused: the GUPPI spectrometer gulp as ONE
// kernel (B200 extension; no single reference counterpart).
//
// Equivalent reference chain (testbench/gpuspec_simple.py:44-55, README):
//   transpose([time,pol,freq,fine_time])           bfTranspose
//   fft(fine_time, apply_fftshift=True)            bfFft  (ci8 -> cf32, x/128)
//   detect('stokes')                               bfMap  (I,Q,U,V)
//   merge_axes(freq, fine_freq); reduce(freq, f)   bfReduce (sum of f bins)
//   accumulate(nframe)                             bfMap  (b = beta*b + a)
// Run unfused, every arrow is an HBM round trip of a tensor up to 8x the input
// (~20x the input bytes per gulp).  Here one CTA owns one coarse
// channel: per frame it reads the 16 KB of interleaved ci8 voltages once,
// runs both polarisations' 4096-point FFTs in registers/shared memory
// (radix-16 Stockham, same butterflies as fft.cu), forms Stokes parameters,
// sums f adjacent fine channels with warp shuffles in the reference's
// left-to-right order, accumulates frames in registers and writes
// 4 x nfft/f floats at the end.  HBM traffic = input + 1/(f*nframe) of it.
// The kernel is fp32-issue bound, not HBM bound.
#include "core.hpp"

namespace bfb {

// --- radix-16 butterfly (same algorithm as fft.cu, float only) ---------------
struct C16 {
	static __host__ __device__ constexpr float c(int k) {
		return k == 0 ? 1.0f : k == 1 ? 0.92387953251128673848f :
		       k == 2 ? 0.70710678118654752440f : k == 3 ? 0.38268343236508977173f :
		       k == 4 ? 0.0f : k == 5 ? -0.38268343236508977173f :
		       k == 6 ? -0.70710678118654752440f : -0.92387953251128673848f;
	}
	static __host__ __device__ constexpr float s(int k) {
		return k == 0 ? 0.0f : k == 1 ? 0.38268343236508977173f :
		       k == 2 ? 0.70710678118654752440f : k == 3 ? 0.92387953251128673848f :
		       k == 4 ? 1.0f : k == 5 ? 0.92387953251128673848f :
		       k == 6 ? 0.70710678118654752440f : 0.38268343236508977173f;
	}
};

template<int R> struct SDft {
	static __device__ __forceinline__ void apply(float* re, float* im) {
		float er[R/2], ei[R/2], qr[R/2], qi[R/2];
#pragma unroll
		for( int k=0; k<R/2; ++k ) { er[k] = re[2*k]; ei[k] = im[2*k]; qr[k] = re[2*k+1]; qi[k] = im[2*k+1]; }
		SDft<R/2>::apply(er, ei);
		SDft<R/2>::apply(qr, qi);
#pragma unroll
		for( int k=0; k<R/2; ++k ) {
			float tr, ti;
			if( k == 0 )        { tr = qr[k]; ti = qi[k]; }
			else if( 4*k == R ) { tr = qi[k]; ti = -qr[k]; }
			else {
				const float c = C16::c(k * (16 / R)), s = C16::s(k * (16 / R));
				tr = qr[k] * c + qi[k] * s;
				ti = qi[k] * c - qr[k] * s;
			}
			re[k]       = er[k] + tr;  im[k]       = ei[k] + ti;
			re[k + R/2] = er[k] - tr;  im[k + R/2] = ei[k] - ti;
		}
	}
};
template<> struct SDft<1> { static __device__ __forceinline__ void apply(float*, float*) {} };

// Shared-memory layout: interleaved complex (one 64-bit access per point) with
// one pad slot per 16 points, which keeps every access pattern of the three
// stages (stride-16 scatter, unit-stride gather, 16-blocked scatter) free of
// bank conflicts within a half-warp.
__device__ __forceinline__ int spad(int i) { return i + (i >> 4); }
enum { SPEC_N = 4096, SPEC_PITCH = SPEC_N + (SPEC_N >> 4) + 1 };

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
	return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// v[t] *= w^t, t = 1..15, from the four exactly tabulated powers w, w^2, w^4,
// w^8: every factor is a product of at most three table values, so the
// twiddles stay within ~2 ulp without fifteen table look-ups.
__device__ __forceinline__ void apply_twiddles(float (&vr)[16], float (&vi)[16],
                                               float2 w1, float2 w2, float2 w4, float2 w8) {
	float2 w[16];
	w[1] = w1; w[2] = w2; w[4] = w4; w[8] = w8;
	w[3] = cmulf(w1, w2);   w[5] = cmulf(w1, w4);   w[6] = cmulf(w2, w4);   w[7] = cmulf(w[3], w4);
	w[9] = cmulf(w1, w8);   w[10] = cmulf(w2, w8);  w[11] = cmulf(w[3], w8); w[12] = cmulf(w4, w8);
	w[13] = cmulf(w[5], w8); w[14] = cmulf(w[6], w8); w[15] = cmulf(w[7], w8);
#pragma unroll
	for( int t=1; t<16; ++t ) {
		float tr = vr[t] * w[t].x - vi[t] * w[t].y;
		vi[t] = vr[t] * w[t].y + vi[t] * w[t].x;
		vr[t] = tr;
	}
}

// 4096-point forward FFT across 256 threads; thread p enters with
// x[p + 256 m] in v[m] and leaves with X[p + 256 t] in v[t].
// tb: per-thread twiddle bases, tb[i*256 + p] = W_256^((p&15) 2^i) for i < 4,
//     W_4096^(p 2^(i-4)) for i >= 4 (lane-contiguous: conflict-free loads).
__device__ __forceinline__ void fft4096(float (&vr)[16], float (&vi)[16], float2* buf,
                                        const float2* __restrict__ tb, int p) {
	// stage 1 (Ns = 1)
	SDft<16>::apply(vr, vi);
	__syncthreads();                              // previous readers of the buffer are done
#pragma unroll
	for( int t=0; t<16; ++t ) buf[spad(16 * p + t)] = make_float2(vr[t], vi[t]);
	__syncthreads();
	// stage 2 (Ns = 16)
#pragma unroll
	for( int m=0; m<16; ++m ) { float2 v = buf[spad(p + 256 * m)]; vr[m] = v.x; vi[m] = v.y; }
	__syncthreads();
	{
		apply_twiddles(vr, vi, tb[p], tb[256 + p], tb[512 + p], tb[768 + p]);
		SDft<16>::apply(vr, vi);
		const int k = p & 15;
		const int j0 = (p - k) * 16 + k;
#pragma unroll
		for( int t=0; t<16; ++t ) buf[spad(j0 + 16 * t)] = make_float2(vr[t], vi[t]);
	}
	__syncthreads();
	// stage 3 (Ns = 256)
#pragma unroll
	for( int m=0; m<16; ++m ) { float2 v = buf[spad(p + 256 * m)]; vr[m] = v.x; vi[m] = v.y; }
	apply_twiddles(vr, vi, tb[1024 + p], tb[1280 + p], tb[1536 + p], tb[1792 + p]);
	SDft<16>::apply(vr, vi);
}

struct SpecParams {
	const char* in;          // ci8 [nframe][nchan][nfft][npol=2]
	long  frame_stride;      // bytes
	long  chan_stride;       // bytes
	float* out;              // f32 [4][nchan * nfft / f_avg]
	long  stokes_stride;     // elements
	int   nframe, nchan;
	int   f_avg;             // power of two, 1..32
	float beta;
	const float2* twiddle;   // 4096 entries exp(-2 pi i k / 4096)
};

// Left-to-right sum of the F values held by F adjacent lanes, returned in all of them.
template<int F>
__device__ __forceinline__ float group_sum(float v, int lane) {
	const int base = lane & ~(F - 1);
	float acc = __shfl_sync(0xffffffffu, v, base);
#pragma unroll
	for( int j=1; j<F; ++j ) acc += __shfl_sync(0xffffffffu, v, base + j);
	return acc;
}

template<int F>
__global__ void __launch_bounds__(256, 2)
spectrometer_kernel(const __grid_constant__ SpecParams P) {
	extern __shared__ __align__(16) unsigned char spec_smem[];
	constexpr int N = SPEC_N;
	float2* buf   = (float2*)spec_smem;              // FFT work buffer
	float2* stash = buf + SPEC_PITCH;                // pol-X spectrum, [t][p]
	float2* tb    = stash + N;                       // twiddle bases, [8][256]
	const int p = threadIdx.x, lane = p & 31;
	const int chan = blockIdx.x;
#pragma unroll
	for( int i=0; i<4; ++i ) {
		tb[i * 256 + p]        = P.twiddle[(((p & 15) << i) * 16) & 4095];   // W_256^((p&15) 2^i)
		tb[(4 + i) * 256 + p]  = P.twiddle[(p << i) & 4095];                 // W_4096^(p 2^i)
	}
	// (the first __syncthreads inside fft4096 orders these writes)

	// Lane j of each F-group keeps the sums of the legs t with t % F == j, so
	// every lane owns 16/F legs per Stokes parameter (F = 32: lanes 0..15 own one).
	constexpr int NACC = (16 + F - 1) / F;
	float acc[4][NACC];
	const int nout = N / F;                       // fine channels after averaging
	const int myslot = lane & (F - 1);
#pragma unroll
	for( int s=0; s<4; ++s )
#pragma unroll
		for( int a=0; a<NACC; ++a ) acc[s][a] = 0.f;
	if( P.beta != 0.f ) {
#pragma unroll
		for( int a=0; a<NACC; ++a ) {
			int t = (F > 16) ? myslot : a * F + myslot;
			if( t < 16 ) {
				int o = (p + 256 * t) / F;
#pragma unroll
				for( int s=0; s<4; ++s )
					acc[s][a] = P.beta * P.out[(long)s * P.stokes_stride + (long)chan * nout + o];
			}
		}
	}
	const char* base = P.in + (long)chan * P.chan_stride;
	// x/128, and (-1)^i for the fftshift: i = p + 256 m has the parity of p
	const float sc = (p & 1) ? -(1.f / 128) : (1.f / 128);
	for( int f=0; f<P.nframe; ++f ) {
		const char4* x = (const char4*)(base + (long)f * P.frame_stride);
		char4 raw[16];
#pragma unroll
		for( int m=0; m<16; ++m ) raw[m] = x[p + 256 * m];
		float vr[16], vi[16];
#pragma unroll
		for( int m=0; m<16; ++m ) { vr[m] = raw[m].x * sc; vi[m] = raw[m].y * sc; }
		fft4096(vr, vi, buf, tb, p);
		// park the X spectrum (only this thread reads it back: no barrier needed)
#pragma unroll
		for( int t=0; t<16; ++t ) stash[t * 256 + p] = make_float2(vr[t], vi[t]);
#pragma unroll
		for( int m=0; m<16; ++m ) { vr[m] = raw[m].z * sc; vi[m] = raw[m].w * sc; }
		fft4096(vr, vi, buf, tb, p);
		// Stokes (blocks/detect.py:102-114 with Complex.hpp arithmetic), then the
		// f_avg sum across adjacent lanes (output index p + 256 t: neighbours in p)
#pragma unroll
		for( int t=0; t<16; ++t ) {
			const float2 X = stash[t * 256 + p];
			const float yr = vr[t], yi = vi[t];
			float xx = X.x * X.x; xx += X.y * X.y;
			float yy = yr * yr;   yy += yi * yi;
			float re = X.x * yr;  re -= X.y * (-yi);
			float im = X.y * yr;  im += X.x * (-yi);
			float sI = group_sum<F>(xx + yy, lane);
			float sQ = group_sum<F>(xx - yy, lane);
			float sU = group_sum<F>(2 * re, lane);
			float sV = group_sum<F>(-2 * im, lane);
			if( F <= 16 ) {
				if( (t % F) == myslot ) {
					const int a = t / F;
					acc[0][a] += sI; acc[1][a] += sQ; acc[2][a] += sU; acc[3][a] += sV;
				}
			} else if( t == myslot ) {
				acc[0][0] += sI; acc[1][0] += sQ; acc[2][0] += sU; acc[3][0] += sV;
			}
		}
	}
#pragma unroll
	for( int a=0; a<NACC; ++a ) {
		int t = (F > 16) ? myslot : a * F + myslot;
		if( t < 16 ) {
			int o = (p + 256 * t) / F;
#pragma unroll
			for( int s=0; s<4; ++s )
				P.out[(long)s * P.stokes_stride + (long)chan * nout + o] = acc[s][a];
		}
	}
}

} // namespace bfb

using namespace bfb;

namespace {
float2* g_twiddle4096[64] = {nullptr};    // per device
BFstatus get_twiddle(const float2** out) {
	int dev = 0;
	BFB_CUDA(cudaGetDevice(&dev), BF_STATUS_DEVICE_ERROR);
	BFB_ASSERT(dev >= 0 && dev < 64, BF_STATUS_INTERNAL_ERROR);
	if( !g_twiddle4096[dev] ) {
		static float2 host[4096];
		const double two_pi = 6.283185307179586476925286766559;
		for( int k=0; k<4096; ++k ) {
			host[k].x = (float)cos(two_pi * k / 4096);
			host[k].y = (float)(-sin(two_pi * k / 4096));
		}
		float2* d = nullptr;
		BFB_CUDA(cudaMalloc((void**)&d, sizeof(host)), BF_STATUS_MEM_ALLOC_FAILED);
		BFB_CUDA(cudaMemcpy(d, host, sizeof(host), cudaMemcpyHostToDevice), BF_STATUS_MEM_OP_FAILED);
		g_twiddle4096[dev] = d;
	}
	*out = g_twiddle4096[dev];
	return BF_STATUS_SUCCESS;
}
} // namespace

extern "C"
BFstatus bfSpectrometerFused(BFarray const* in, BFarray const* out, int nfft, int f_avg, double beta) {
	BFB_ASSERT(in && out, BF_STATUS_INVALID_POINTER);
	BFB_ASSERT(space_on_device(in->space) && space_on_device(out->space), BF_STATUS_UNSUPPORTED_SPACE);
	BFB_ASSERT(in->dtype == BF_DTYPE_CI8,  BF_STATUS_UNSUPPORTED_DTYPE);
	BFB_ASSERT(out->dtype == BF_DTYPE_F32, BF_STATUS_UNSUPPORTED_DTYPE);
	BFB_ASSERT(nfft == 4096, BF_STATUS_UNSUPPORTED_SHAPE);
	BFB_ASSERT(f_avg >= 1 && f_avg <= 32 && (f_avg & (f_avg - 1)) == 0, BF_STATUS_UNSUPPORTED_SHAPE);
	// in: [nframe, nchan, nfft, 2] (3-D [nchan, nfft, 2] means one frame)
	BFB_ASSERT(in->ndim == 3 || in->ndim == 4, BF_STATUS_INVALID_SHAPE);
	int o = in->ndim - 3;
	long nframe = o ? in->shape[0] : 1, nchan = in->shape[o];
	BFB_ASSERT(in->shape[o+1] == nfft && in->shape[o+2] == 2, BF_STATUS_INVALID_SHAPE);
	BFB_ASSERT(in->strides[o+2] == 2 && in->strides[o+1] == 4, BF_STATUS_UNSUPPORTED_STRIDE);
	long frame_stride = o ? in->strides[0] : 0, chan_stride = in->strides[o];
	BFB_ASSERT((uintptr_t)in->data % 4 == 0 && frame_stride % 4 == 0 && chan_stride % 4 == 0,
	           BF_STATUS_UNSUPPORTED_STRIDE);
	// out: [4, nchan*nfft/f_avg] (leading unit dims allowed)
	int od = out->ndim;
	BFB_ASSERT(od >= 2, BF_STATUS_INVALID_SHAPE);
	for( int d=0; d<od-2; ++d ) BFB_ASSERT(out->shape[d] == 1, BF_STATUS_INVALID_SHAPE);
	BFB_ASSERT(out->shape[od-2] == 4 && out->shape[od-1] == nchan * nfft / f_avg, BF_STATUS_INVALID_SHAPE);
	BFB_ASSERT(out->strides[od-1] == 4 && out->strides[od-2] % 4 == 0, BF_STATUS_UNSUPPORTED_STRIDE);
	if( nchan == 0 || nframe == 0 ) return BF_STATUS_SUCCESS;
	SpecParams P;
	P.in = (const char*)in->data; P.frame_stride = frame_stride; P.chan_stride = chan_stride;
	P.out = (float*)out->data; P.stokes_stride = out->strides[od-2] / 4;
	P.nframe = (int)nframe; P.nchan = (int)nchan; P.f_avg = f_avg; P.beta = (float)beta;
	BFstatus s = get_twiddle(&P.twiddle);
	if( s != BF_STATUS_SUCCESS ) return s;
	size_t smem = ((size_t)SPEC_PITCH + SPEC_N + 8 * 256) * sizeof(float2);
	cudaStream_t st = thread_stream();
#define BFB_SPEC(F_) do { \
		BFB_CUDA(cudaFuncSetAttribute(spectrometer_kernel<F_>, \
			cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), BF_STATUS_INTERNAL_ERROR); \
		spectrometer_kernel<F_><<<(unsigned)nchan, 256, smem, st>>>(P); } while(0)
	switch( f_avg ) {
	case  1: BFB_SPEC(1);  break;
	case  2: BFB_SPEC(2);  break;
	case  4: BFB_SPEC(4);  break;
	case  8: BFB_SPEC(8);  break;
	case 16: BFB_SPEC(16); break;
	default: BFB_SPEC(32); break;
	}
#undef BFB_SPEC
	count_launch();
	BFB_CUDA(cudaGetLastError(), BF_STATUS_INTERNAL_ERROR);
	return BF_STATUS_SUCCESS;
}
