// linalg.cu -- bfLinAlg* for sm_100a: the cross-correlation ("X-engine") call
// behind bf.blocks.correlate.
//
// Replaces: src/linalg.cu:877-904 (bfLinAlgMatMul), :242-357 (bfMatMul_aa),
// :190-226 (dispatch to the xGPU-style kernel) and
// src/linalg_kernels.cu:154-643 (bf_cherk_N diagonal/off-diagonal kernels).
//
// Semantics kept: with u = b (b^H.b form) or u(t,i) = conj(a[i][t]) (a.a^H form)
//   c[..., i, j] = alpha * sum_t conj(u[t][i]) * u[t][j] + beta * c[..., i, j]   for i >= j
// and every element above the diagonal is left untouched
// (python/bifrost/blocks/correlate.py:69, test/test_pipeline.py:258-298,
// test/test_linalg.py:168-185).  Integer inputs accumulate exactly in int32.
//
// Two kernels:
//  * corr_tc_kernel -- tcgen05 int8 tensor cores.  The interleaved complex
//    samples of one frequency channel form a real row-major matrix
//    M[t][2i+p] (p = re/im).  Its Gram matrix G = M^T M holds all four real
//    products of every pair: Re C_ij = G[2i][2j] + G[2i+1][2j+1],
//    Im C_ij = G[2i][2j+1] - G[2i+1][2j].  Time (the reduction dim) is the slow
//    axis in memory, i.e. both MMA operands are MN-major: TMA drops
//    [64 t][128 B] boxes with the 128-byte swizzle straight into the canonical
//    MN-major UMMA layout, one elected thread issues
//    tcgen05.mma.kind::i8 (M=128, N=128, K=32) into a TMEM accumulator, and four
//    epilogue warps pull the int32 tile out of TMEM, pair rows with a lane
//    shuffle, and write the lower-triangular cf32 tile through shared memory
//    so that global stores are row-contiguous.  Only tiles on or below the
//    diagonal are launched; diagonal tiles load one operand.
//  * corr_simt_kernel -- shared-memory tiled SIMT fallback for layouts TMA
//    cannot describe (strides not multiples of 16 B, a.a^H form, ci16/cf32).
#include "core.hpp"
#include "shape.hpp"

#include <cuda.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace bfb {

// =============================================================== SIMT fallback
struct SimtParams {
	const char* u;           // element (t, i) at u + t*stride_t + i*stride_i (bytes)
	long  stride_t, stride_i, stride_b;
	float2* c; long c_row, c_batch;      // float2 units
	int   n, ntime, nbatch;
	float alpha, beta;
	int   conj_u;            // u = conj(memory)
};

template<typename I> struct CIn { I x, y; };

template<typename I, typename Acc>
__global__ void __launch_bounds__(256)
corr_simt_kernel(SimtParams P) {
	// 32x32 output tile per CTA, 16x16 threads, 2x2 outputs per thread
	__shared__ float2 su_i[32][33];
	__shared__ float2 su_j[32][33];
	const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
	const int ntile = (P.n + 31) / 32;
	int tile = blockIdx.x;
	// lower-triangular tile index -> (ti, tj), ti >= tj
	int ti = (int)((sqrtf(8.f * tile + 1.f) - 1.f) * 0.5f);
	while( ti * (ti + 1) / 2 > tile ) --ti;
	while( (ti + 1) * (ti + 2) / 2 <= tile ) ++ti;
	int tj = tile - ti * (ti + 1) / 2;
	(void)ntile;
	const int b = blockIdx.y;
	const char* ub = P.u + (long)b * P.stride_b;
	Acc re[2][2] = {{0, 0}, {0, 0}}, im[2][2] = {{0, 0}, {0, 0}};
	for( int t0 = 0; t0 < P.ntime; t0 += 32 ) {
		// stage 32 time samples of the i-block and the j-block
		for( int k = threadIdx.x; k < 32 * 32; k += 256 ) {
			int tt = k >> 5, ii = k & 31;
			int t = t0 + tt;
			float2 vi = make_float2(0.f, 0.f), vj = make_float2(0.f, 0.f);
			if( t < P.ntime ) {
				int gi = ti * 32 + ii, gj = tj * 32 + ii;
				if( gi < P.n ) {
					CIn<I> v = *(const CIn<I>*)(ub + (long)t * P.stride_t + (long)gi * P.stride_i);
					vi = make_float2((float)v.x, (float)v.y);
				}
				if( gj < P.n ) {
					CIn<I> v = *(const CIn<I>*)(ub + (long)t * P.stride_t + (long)gj * P.stride_i);
					vj = make_float2((float)v.x, (float)v.y);
				}
			}
			su_i[tt][ii] = vi; su_j[tt][ii] = vj;
		}
		__syncthreads();
#pragma unroll 4
		for( int tt = 0; tt < 32; ++tt ) {
			float2 a[2] = {su_i[tt][ty], su_i[tt][ty + 16]};
			float2 c[2] = {su_j[tt][tx], su_j[tt][tx + 16]};
#pragma unroll
			for( int p=0; p<2; ++p )
#pragma unroll
				for( int q=0; q<2; ++q ) {
					// conj(u_i) * u_j
					re[p][q] += (Acc)a[p].x * (Acc)c[q].x + (Acc)a[p].y * (Acc)c[q].y;
					im[p][q] += (Acc)a[p].x * (Acc)c[q].y - (Acc)a[p].y * (Acc)c[q].x;
				}
		}
		__syncthreads();
	}
	float2* cb = P.c + (long)b * P.c_batch;
#pragma unroll
	for( int p=0; p<2; ++p )
#pragma unroll
		for( int q=0; q<2; ++q ) {
			int i = ti * 32 + ty + 16 * p, j = tj * 32 + tx + 16 * q;
			if( i < P.n && j < P.n && i >= j ) {
				float vr = (float)re[p][q], vi = (float)im[p][q];
				if( P.conj_u ) vi = -vi;
				float2* dst = cb + (long)i * P.c_row + j;
				float2 o = make_float2(P.alpha * vr, P.alpha * vi);
				if( P.beta != 0.f ) { float2 old = *dst; o.x += P.beta * old.x; o.y += P.beta * old.y; }
				*dst = o;
			}
		}
}

// ============================================================ tcgen05 kernel
enum { TC_KT = 64, TC2_STAGES = 3, TC_TILE_BYTES = TC_KT * 128, TC_THREADS = 192,
       TC_STAGE_BYTES = 3 * TC_TILE_BYTES };    // A | B0 | B1

struct TcParams {
	float2* c; long c_row, c_batch;      // float2 units
	int   n;                 // complex elements per time sample
	int   ntime;
	int   ntile;             // lower-triangular 128x128 tiles per batch
	float alpha, beta;
	int   conj_u;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
	return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"WAIT_LOOP:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra DONE;\n\t"
		"bra WAIT_LOOP;\n\t"
		"DONE:\n\t}"
		:: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int x, int y, int z) {
	asm volatile(
		"cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
		" [%0], [%1, {%3, %4, %5}], [%2];"
		:: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
// 64-bit shared-memory matrix descriptor: MN-major, 128-byte swizzle, one
// 128-byte atom wide; 8-row groups are 1024 B apart (SBO).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr) {
	uint64_t d = 0;
	d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address
	d |= (uint64_t)(TC_TILE_BYTES >> 4) << 16;        // leading byte offset: next 128-byte atom along MN (N = 256)
	d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset between 8-row groups
	d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
	d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
	return d;
}
// Instruction descriptor: S32 accumulate, signed int8 A and B, both MN-major,
// M = 128, N = 128 or 256.
__device__ __forceinline__ uint32_t umma_idesc_i8(uint32_t n_dim) {
	uint32_t d = 0;
	d |= 2u << 4;            // c_format = S32
	d |= 1u << 7;            // a_format = INT8 (signed)
	d |= 1u << 10;           // b_format = INT8 (signed)
	d |= 1u << 15;           // a_major = MN
	d |= 1u << 16;           // b_major = MN
	d |= (n_dim >> 3) << 17; // n_dim
	d |= (128u >> 4) << 24;  // m_dim
	return d;
}

// Epilogue of one (row block I, column blocks J0 .. J0+nb-1) unit, run by warps
// 2..5: TMEM -> registers (32x32b.x32), combine the interleaved re/im rows of
// the real Gram matrix into complex visibilities with a lane-pair shuffle,
// stage through shared memory and store coalesced lower-triangle cf32.
__device__ __forceinline__ void tc_epilogue(uint64_t* tmem_bar, uint32_t tmem_base, float2* staging,
                                            TcParams const& P, int batch, int I, int J0, int nb,
                                            int warp, int lane) {
	mbar_wait(tmem_bar, 0);
	asm volatile("tcgen05.fence::after_thread_sync;");
	const int quarter = warp & 3;                       // TMEM lane quarter this warp may read
	const int m = quarter * 32 + lane;                  // row of the 128 x 256 int32 tile
	const int il = m >> 1, par = m & 1;                 // complex row, re/im row of the pair
	float* stf = (float*)staging;
	const int te = threadIdx.x - 64;                    // 0..127
	float2* cb = P.c + (long)batch * P.c_batch;
	for( int h=0; h<nb; ++h ) {
		const int J = J0 + h;
#pragma unroll 1
		for( int c0=0; c0<128; c0+=32 ) {
			uint32_t r[32];
			const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(128 * h + c0);
			asm volatile(
				"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
				"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
				"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
				: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
				  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
				  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
				  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
				: "r"(taddr));
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
			for( int q=0; q<16; ++q ) {
				// even row 2i holds (G[2i][2j], G[2i][2j+1]); odd row holds
				// (G[2i+1][2j], G[2i+1][2j+1]); swap the second entries.
				int mine0 = (int)r[2*q], mine1 = (int)r[2*q+1];
				int other1 = __shfl_xor_sync(0xffffffffu, mine1, 1);
				// even lane: Re = G[2i][2j] + G[2i+1][2j+1]; odd lane: Im = G[2i][2j+1] - G[2i+1][2j]
				int v = par ? (other1 - mine0) : (mine0 + other1);
				int jl = (c0 >> 1) + q;
				stf[(il * 65 + jl) * 2 + par] = (float)v;
			}
		}
		// all four epilogue warps have staged their rows
		asm volatile("bar.sync 1, 128;" ::: "memory");
		for( int idx = te; idx < 64 * 64; idx += 128 ) {
			int il2 = idx >> 6, jl2 = idx & 63;
			int i = I * 64 + il2, j = J * 64 + jl2;
			if( i < P.n && j < P.n && i >= j ) {
				float2 v = staging[il2 * 65 + jl2];
				if( P.conj_u ) v.y = -v.y;
				float2 o = make_float2(P.alpha * v.x, P.alpha * v.y);
				float2* dst = cb + (long)i * P.c_row + j;
				if( P.beta != 0.f ) { float2 old = *dst; o.x += P.beta * old.x; o.y += P.beta * old.y; }
				*dst = o;
			}
		}
		// the staging buffer is reused by the second column block
		asm volatile("bar.sync 1, 128;" ::: "memory");
	}
	}

template<int TC_STAGES>
__global__ void __launch_bounds__(TC_THREADS)
corr_tc_kernel(const __grid_constant__ CUtensorMap tmap, TcParams P) {
	extern __shared__ __align__(1024) unsigned char tc_smem[];
	// layout: [stages][A 8 KB | B0 8 KB | B1 8 KB] | staging float2[64][65] | barriers
	// the swizzled operand tiles need 1024-byte alignment in the shared window
	unsigned char* tiles = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);
	float2* staging = (float2*)(tiles + TC_STAGES * TC_STAGE_BYTES);
	uint64_t* full_bar  = (uint64_t*)(staging + 64 * 65);
	uint64_t* empty_bar = full_bar + TC_STAGES;
	uint64_t* tmem_bar  = empty_bar + TC_STAGES;
	uint32_t* tmem_slot = (uint32_t*)(tmem_bar + 1);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int batch = blockIdx.y;
	// Work unit: row block I (128 rows of the real Gram matrix) x a pair of
	// column blocks (2 JJ, 2 JJ + 1) with 2 JJ <= I.  The second block is
	// dropped when it lies above the diagonal (nb = 1).  A 128 x 256 tile moves
	// 24 KB per K slab for two blocks instead of 32 KB: the kernel is bound by
	// L2 -> SM operand traffic, not by the tensor pipe.
	int I = 0, rem = blockIdx.x;
	while( rem >= I / 2 + 1 ) { rem -= I / 2 + 1; ++I; }
	const int J0 = 2 * rem;
	const int nb = (J0 + 1 <= I) ? 2 : 1;
	// operand aliasing: a column block equal to the row block is loaded once
	const bool a_is_b0 = (J0 == I), a_is_b1 = (nb == 2 && J0 + 1 == I);
	const int nk = (P.ntime + TC_KT - 1) / TC_KT;

	if( threadIdx.x == 0 ) {
		for( int s=0; s<TC_STAGES; ++s ) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
		mbar_init(tmem_bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if( warp == 1 ) {
		// 256 TMEM columns: a 128 x 256 int32 accumulator; the address lands in smem
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
		             :: "r"(smem_u32(tmem_slot)), "n"(256));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tmem_base = *tmem_slot;

	if( warp == 0 ) {
		// ===================== TMA producer =====================
		if( lane == 0 ) {
			const uint32_t bytes = (uint32_t)TC_TILE_BYTES * (nb + ((a_is_b0 || a_is_b1) ? 0 : 1));
			for( int it=0; it<nk; ++it ) {
				const int s = it % TC_STAGES;
				const uint32_t ph = (it / TC_STAGES) & 1;
				mbar_wait(&empty_bar[s], ph ^ 1);
				unsigned char* a = tiles + (size_t)s * TC_STAGE_BYTES;
				mbar_expect_tx(&full_bar[s], bytes);
				if( !(a_is_b0 || a_is_b1) ) tma_load_3d(a, &tmap, &full_bar[s], 128 * I, it * TC_KT, batch);
				tma_load_3d(a + TC_TILE_BYTES, &tmap, &full_bar[s], 128 * J0, it * TC_KT, batch);
				if( nb == 2 ) tma_load_3d(a + 2 * TC_TILE_BYTES, &tmap, &full_bar[s], 128 * (J0 + 1), it * TC_KT, batch);
			}
		}
	} else if( warp == 1 ) {
		// ===================== MMA issuer =====================
		const uint32_t idesc = umma_idesc_i8(nb == 2 ? 256u : 128u);
		for( int it=0; it<nk; ++it ) {
			const int s = it % TC_STAGES;
			const uint32_t ph = (it / TC_STAGES) & 1;
			mbar_wait(&full_bar[s], ph);
			asm volatile("tcgen05.fence::after_thread_sync;");
			if( lane == 0 ) {
				const uint32_t st_addr = smem_u32(tiles + (size_t)s * TC_STAGE_BYTES);
				const uint32_t b_addr = st_addr + TC_TILE_BYTES;
				const uint32_t a_addr = a_is_b0 ? b_addr : (a_is_b1 ? b_addr + TC_TILE_BYTES : st_addr);
#pragma unroll
				for( int kk=0; kk<TC_KT/32; ++kk ) {
					const uint64_t da = umma_desc_mn_sw128(a_addr + kk * 32 * 128);
					const uint64_t db = umma_desc_mn_sw128(b_addr + kk * 32 * 128);
					const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
					asm volatile(
						"{\n\t.reg .pred p;\n\t"
						"setp.ne.b32 p, %4, 0;\n\t"
						"tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
						:: "r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
				}
				// release the smem stage once these MMAs have consumed it
				asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
				             :: "r"(smem_u32(&empty_bar[s])) : "memory");
				if( it == nk - 1 ) {
					asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
					             :: "r"(smem_u32(tmem_bar)) : "memory");
				}
			}
			__syncwarp();
		}
	} else {
		tc_epilogue(tmem_bar, tmem_base, staging, P, batch, I, J0, nb, warp, lane);
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if( warp == 1 ) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(256));
	}
}

// ---------------------------------------------------------------------------
// Cluster variant: a CTA pair owns the 256 x 256 super tile (row blocks 2 II
// and 2 II + 1) x (column blocks 2 JJ and 2 JJ + 1), JJ <= II.  Each CTA loads
// ONE of the two column blocks per K slab and TMA-multicasts it to both CTAs, so
// the pair moves A0 + A1 + B0 + B1 = 32 KB per slab for four output blocks
// (8 KB per block instead of 12).  On diagonal super tiles the row blocks ARE
// the column blocks and nothing else is loaded.  A stage may be refilled only
// after BOTH CTAs' MMAs have released it: every tcgen05.commit is multicast to
// the empty barrier of both CTAs (count 2).
__device__ __forceinline__ uint32_t cluster_ctarank() {
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(void* dst, const CUtensorMap* map, uint64_t* bar,
                                               int x, int y, int z, uint16_t mask) {
	asm volatile(
		"cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
		" [%0], [%1, {%3, %4, %5}], [%2], %6;"
		:: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS)
corr_tc2_kernel(const __grid_constant__ CUtensorMap tmap, TcParams P) {
	extern __shared__ __align__(1024) unsigned char tc_smem[];
	unsigned char* tiles = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);
	float2* staging = (float2*)(tiles + TC2_STAGES * TC_STAGE_BYTES);
	uint64_t* full_bar  = (uint64_t*)(staging + 64 * 65);
	uint64_t* empty_bar = full_bar + TC2_STAGES;
	uint64_t* tmem_bar  = empty_bar + TC2_STAGES;
	uint32_t* tmem_slot = (uint32_t*)(tmem_bar + 1);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int batch = blockIdx.y;
	const int rank = (int)cluster_ctarank();            // 0 / 1: upper / lower row block of the pair
	// super tile index -> (II, JJ), II >= JJ
	const int st = blockIdx.x >> 1;
	int II = (int)((sqrtf(8.f * st + 1.f) - 1.f) * 0.5f);
	while( II * (II + 1) / 2 > st ) --II;
	while( (II + 1) * (II + 2) / 2 <= st ) ++II;
	const int JJ = st - II * (II + 1) / 2;
	const bool sdiag = (II == JJ);
	const int I = 2 * II + rank, J0 = 2 * JJ;
	// column blocks this CTA needs: on the diagonal super tile the upper row
	// block only has its own diagonal block
	const int nb = (sdiag && rank == 0) ? 1 : 2;
	const int nk = (P.ntime + TC_KT - 1) / TC_KT;

	if( threadIdx.x == 0 ) {
		for( int s=0; s<TC2_STAGES; ++s ) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 2); }
		mbar_init(tmem_bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if( warp == 1 ) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
		             :: "r"(smem_u32(tmem_slot)), "n"(256));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tmem_base = *tmem_slot;
	// the peer's barriers must be initialised before anything is multicast to it
	cluster_sync_all();

	if( warp == 0 ) {
		// ===================== TMA producer =====================
		if( lane == 0 ) {
			// both column blocks always arrive (one from each CTA); the row block
			// is loaded only off the diagonal
			const uint32_t bytes = (uint32_t)TC_TILE_BYTES * (sdiag ? 2 : 3);
			for( int it=0; it<nk; ++it ) {
				const int s = it % TC2_STAGES;
				const uint32_t ph = (it / TC2_STAGES) & 1;
				mbar_wait(&empty_bar[s], ph ^ 1);            // both CTAs have released the stage
				unsigned char* a = tiles + (size_t)s * TC_STAGE_BYTES;
				mbar_expect_tx(&full_bar[s], bytes);
				if( !sdiag ) tma_load_3d(a, &tmap, &full_bar[s], 128 * I, it * TC_KT, batch);
				tma_load_3d_mc(a + (1 + rank) * TC_TILE_BYTES, &tmap, &full_bar[s],
				               128 * (J0 + rank), it * TC_KT, batch, (uint16_t)0x3);
			}
		}
	} else if( warp == 1 ) {
		// ===================== MMA issuer =====================
		const uint32_t idesc = umma_idesc_i8(nb == 2 ? 256u : 128u);
		for( int it=0; it<nk; ++it ) {
			const int s = it % TC2_STAGES;
			const uint32_t ph = (it / TC2_STAGES) & 1;
			mbar_wait(&full_bar[s], ph);
			asm volatile("tcgen05.fence::after_thread_sync;");
			if( lane == 0 ) {
				const uint32_t st_addr = smem_u32(tiles + (size_t)s * TC_STAGE_BYTES);
				const uint32_t b_addr = st_addr + TC_TILE_BYTES;
				const uint32_t a_addr = sdiag ? b_addr + rank * TC_TILE_BYTES : st_addr;
#pragma unroll
				for( int kk=0; kk<TC_KT/32; ++kk ) {
					const uint64_t da = umma_desc_mn_sw128(a_addr + kk * 32 * 128);
					const uint64_t db = umma_desc_mn_sw128(b_addr + kk * 32 * 128);
					const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
					asm volatile(
						"{\n\t.reg .pred p;\n\t"
						"setp.ne.b32 p, %4, 0;\n\t"
						"tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
						:: "r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
				}
				// release the stage in BOTH CTAs once these MMAs have consumed it
				asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
				             :: "r"(smem_u32(&empty_bar[s])), "h"((uint16_t)0x3) : "memory");
				if( it == nk - 1 ) {
					asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
					             :: "r"(smem_u32(tmem_bar)) : "memory");
				}
			}
			__syncwarp();
		}
	} else {
		tc_epilogue(tmem_bar, tmem_base, staging, P, batch, I, J0, nb, warp, lane);
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if( warp == 1 ) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(256));
	}
	// neither CTA may exit while the other can still multicast into it
	cluster_sync_all();
}

// ---------------------------------------------------------------------------
// c = alpha a.b + beta c for ci8 operands on the tensor cores (the reference
// hands this case to cublasCgemmEx with CUDA_C_8I inputs, src/linalg.cu:406-424;
// it is the beamformer's product: weights x voltages).  Same machine as the
// correlator above, with two different operands: with At = a^T (k slow) read as
// the real matrix At[k][2m+p] and b as B[k][2n+q] (p, q = re/im), the real
// int8 product G = At^T B holds the four real products of every (m, n):
//   Re c = G[2m][2n] - G[2m+1][2n+1],   Im c = G[2m][2n+1] + G[2m+1][2n]
// (signs flip with the operands' conjugation flags).  Both operands are
// MN-major exactly like the correlator's, so TMA boxes, UMMA descriptors, the
// stage pipeline and the TMEM epilogue are shared; the sums are exact integers.
// A work unit is a 128 x 256 tile of G = 64 x 128 complex outputs.
// ---------------------------------------------------------------------------
struct AbTcParams {
	float2* c; long c_row, c_batch;      // float2 units
	int   M, N, K;                       // complex rows, complex columns, reduction length
	int   njj;                           // column-block pairs per row block
	int   a_batched;                     // 0: one At for every batch entry
	float alpha, beta;
	float s_re, s_01, s_10;              // Re = G00 + s_re G11, Im = s_01 G01 + s_10 G10
};

__device__ __forceinline__ void ab_epilogue(uint64_t* tmem_bar, uint32_t tmem_base, float2* staging,
                                            AbTcParams const& P, int batch, int I, int J0, int nb,
                                            int warp, int lane) {
	mbar_wait(tmem_bar, 0);
	asm volatile("tcgen05.fence::after_thread_sync;");
	const int quarter = warp & 3;
	const int m = quarter * 32 + lane;
	const int il = m >> 1, par = m & 1;
	float* stf = (float*)staging;
	const int te = threadIdx.x - 64;
	float2* cb = P.c + (long)batch * P.c_batch;
	for( int h=0; h<nb; ++h ) {
		const int J = J0 + h;
#pragma unroll 1
		for( int c0=0; c0<128; c0+=32 ) {
			uint32_t r[32];
			const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(128 * h + c0);
			asm volatile(
				"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
				"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
				"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
				: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
				  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
				  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
				  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
				: "r"(taddr));
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
			for( int q=0; q<16; ++q ) {
				// even row 2m: (G00, G01); odd row 2m+1: (G10, G11); swap the second entries
				const int mine0 = (int)r[2*q], mine1 = (int)r[2*q+1];
				const int other1 = __shfl_xor_sync(0xffffffffu, mine1, 1);
				const float v = par ? (P.s_01 * (float)other1 + P.s_10 * (float)mine0)
				                    : ((float)mine0 + P.s_re * (float)other1);
				const int jl = (c0 >> 1) + q;
				stf[(il * 65 + jl) * 2 + par] = v;
			}
		}
		asm volatile("bar.sync 1, 128;" ::: "memory");
		for( int idx = te; idx < 64 * 64; idx += 128 ) {
			const int il2 = idx >> 6, jl2 = idx & 63;
			const int i = I * 64 + il2, j = J * 64 + jl2;
			if( i < P.M && j < P.N ) {
				const float2 v = staging[il2 * 65 + jl2];
				float2 o = make_float2(P.alpha * v.x, P.alpha * v.y);
				float2* dst = cb + (long)i * P.c_row + j;
				if( P.beta != 0.f ) { float2 old = *dst; o.x += P.beta * old.x; o.y += P.beta * old.y; }
				*dst = o;
			}
		}
		asm volatile("bar.sync 1, 128;" ::: "memory");
	}
}

template<int TC_STAGES>
__global__ void __launch_bounds__(TC_THREADS)
ab_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, AbTcParams P) {
	extern __shared__ __align__(1024) unsigned char tc_smem[];
	unsigned char* tiles = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);
	float2* staging = (float2*)(tiles + TC_STAGES * TC_STAGE_BYTES);
	uint64_t* full_bar  = (uint64_t*)(staging + 64 * 65);
	uint64_t* empty_bar = full_bar + TC_STAGES;
	uint64_t* tmem_bar  = empty_bar + TC_STAGES;
	uint32_t* tmem_slot = (uint32_t*)(tmem_bar + 1);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int batch = blockIdx.y;
	const int I = blockIdx.x / P.njj, J0 = 2 * (blockIdx.x % P.njj);
	const int nb = ((J0 + 1) * 64 < P.N) ? 2 : 1;
	const int nk = (P.K + TC_KT - 1) / TC_KT;

	if( threadIdx.x == 0 ) {
		for( int s=0; s<TC_STAGES; ++s ) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
		mbar_init(tmem_bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if( warp == 1 ) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
		             :: "r"(smem_u32(tmem_slot)), "n"(256));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;");
	const uint32_t tmem_base = *tmem_slot;

	if( warp == 0 ) {
		if( lane == 0 ) {
			const uint32_t bytes = (uint32_t)TC_TILE_BYTES * (1 + nb);
			for( int it=0; it<nk; ++it ) {
				const int s = it % TC_STAGES;
				const uint32_t ph = (it / TC_STAGES) & 1;
				mbar_wait(&empty_bar[s], ph ^ 1);
				unsigned char* a = tiles + (size_t)s * TC_STAGE_BYTES;
				mbar_expect_tx(&full_bar[s], bytes);
				tma_load_3d(a, &tmap_a, &full_bar[s], 128 * I, it * TC_KT, P.a_batched ? batch : 0);
				tma_load_3d(a + TC_TILE_BYTES, &tmap_b, &full_bar[s], 128 * J0, it * TC_KT, batch);
				if( nb == 2 ) tma_load_3d(a + 2 * TC_TILE_BYTES, &tmap_b, &full_bar[s], 128 * (J0 + 1), it * TC_KT, batch);
			}
		}
	} else if( warp == 1 ) {
		const uint32_t idesc = umma_idesc_i8(nb == 2 ? 256u : 128u);
		for( int it=0; it<nk; ++it ) {
			const int s = it % TC_STAGES;
			const uint32_t ph = (it / TC_STAGES) & 1;
			mbar_wait(&full_bar[s], ph);
			asm volatile("tcgen05.fence::after_thread_sync;");
			if( lane == 0 ) {
				const uint32_t a_addr = smem_u32(tiles + (size_t)s * TC_STAGE_BYTES);
				const uint32_t b_addr = a_addr + TC_TILE_BYTES;
#pragma unroll
				for( int kk=0; kk<TC_KT/32; ++kk ) {
					const uint64_t da = umma_desc_mn_sw128(a_addr + kk * 32 * 128);
					const uint64_t db = umma_desc_mn_sw128(b_addr + kk * 32 * 128);
					const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
					asm volatile(
						"{\n\t.reg .pred p;\n\t"
						"setp.ne.b32 p, %4, 0;\n\t"
						"tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
						:: "r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
				}
				asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
				             :: "r"(smem_u32(&empty_bar[s])) : "memory");
				if( it == nk - 1 ) {
					asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
					             :: "r"(smem_u32(tmem_bar)) : "memory");
				}
			}
			__syncwarp();
		}
	} else {
		ab_epilogue(tmem_bar, tmem_base, staging, P, batch, I, J0, nb, warp, lane);
	}
	asm volatile("tcgen05.fence::before_thread_sync;");
	__syncthreads();
	if( warp == 1 ) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(256));
	}
}

// a[m][k] (any strides) -> At[batch][k][m] ci8, row pitch `pitch` bytes
__global__ void __launch_bounds__(256)
ab_transpose_a_kernel(const char* a, long sm, long sk, long sbatch, char* at, long pitch, long at_batch, int M, int K) {
	__shared__ short tile[32][33];
	const int bx = blockIdx.x * 32, by = blockIdx.y * 32;          // bx: k, by: m
	const char* ab = a + (long)blockIdx.z * sbatch;
	char* tb = at + (long)blockIdx.z * at_batch;
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
	for( int r=ty; r<32; r+=8 ) {
		const int m = by + r, k = bx + tx;
		tile[r][tx] = (m < M && k < K) ? *(const short*)(ab + (long)m * sm + (long)k * sk) : (short)0;
	}
	__syncthreads();
	for( int r=ty; r<32; r+=8 ) {
		const int k = bx + r, m = by + tx;
		if( k < K && m < M ) *(short*)(tb + (long)k * pitch + 2L * m) = tile[tx][r];
	}
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
	static PFN_encodeTiled fn = nullptr;
	static bool tried = false;
	if( !tried ) {
		tried = true;
		void* p = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if( cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
		    qres == cudaDriverEntryPointSuccess ) {
			fn = (PFN_encodeTiled)p;
		}
	}
	return fn;
}

} // namespace bfb

using namespace bfb;

struct BFlinalg_impl {
	// a^T staging of the tensor-core a.b path (ci8), grown on demand
	void*  at_buf = nullptr;
	size_t at_size = 0;
	~BFlinalg_impl() { if( at_buf ) cudaFree(at_buf); }
};

namespace {

// c[..,i,j] = alpha sum_t conj(u(t,i)) u(t,j) + beta c, lower triangle.
// u(t,i) lives at data + t*stride_t + i*stride_i (+ batch offsets).
BFstatus correlate(BFdtype utype, const void* udata, long stride_t, long stride_i,
                   int conj_u, long n, long ntime,
                   int nbdim, const long* bshape, const long* ustr, const long* cstr,
                   void* cdata, long c_row_bytes, double alpha, double beta, cudaStream_t st) {
	if( n == 0 ) return BF_STATUS_SUCCESS;
	BFB_ASSERT(c_row_bytes % 8 == 0, BF_STATUS_UNSUPPORTED_STRIDE);
	// Kernel batch = innermost batch dim; outer batch dims are looped on the host.
	long kb = 1, kb_ustr = 0, kb_cstr = 0;
	int kdim = -1;
	for( int d=0; d<nbdim; ++d ) if( bshape[d] >= kb ) { kb = bshape[d]; kdim = d; }
	if( kdim >= 0 ) { kb_ustr = ustr[kdim]; kb_cstr = cstr[kdim]; BFB_ASSERT(kb_cstr % 8 == 0, BF_STATUS_UNSUPPORTED_STRIDE); }
	BFB_ASSERT(kb <= 65535, BF_STATUS_UNSUPPORTED_SHAPE);
	long nouter = 1;
	for( int d=0; d<nbdim; ++d ) if( d != kdim ) nouter *= bshape[d];
	bool tc_ok = utype == BF_DTYPE_CI8 && stride_i == 2 && stride_t % 16 == 0 &&
	             (kb == 1 || (kb_ustr % 16 == 0 && kb_ustr >= 16)) && n >= 16 && ntime >= 1 &&
	             ntime * 2L * 127 * 127 < (1L << 31) && get_encode_fn() != nullptr &&
	             getenv("BFB_LINALG_SIMT") == nullptr;
	for( long o=0; o<nouter; ++o ) {
		long rem = o, uoff = 0, coff = 0;
		for( int d=nbdim-1; d>=0; --d ) {
			if( d == kdim ) continue;
			long r = rem % bshape[d]; rem /= bshape[d];
			uoff += r * ustr[d]; coff += r * cstr[d];
		}
		const char* ub = (const char*)udata + uoff;
		float2* cb = (float2*)((char*)cdata + coff);
		bool use_tc = tc_ok && ((uintptr_t)ub % 16 == 0);
		if( use_tc ) {
			CUtensorMap tmap;
			cuuint64_t gdim[3] = {(cuuint64_t)(2 * n), (cuuint64_t)ntime, (cuuint64_t)kb};
			cuuint64_t gstr[2] = {(cuuint64_t)stride_t, (cuuint64_t)(kb > 1 ? kb_ustr : stride_t * ntime)};
			if( gstr[1] % 16 ) gstr[1] = round_up<cuuint64_t>(gstr[1], 16);
			cuuint32_t box[3] = {128, TC_KT, 1};
			cuuint32_t estr[3] = {1, 1, 1};
			CUresult res = get_encode_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)ub, gdim, gstr,
			                               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
			                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
			if( res != CUDA_SUCCESS ) use_tc = false;
			if( use_tc ) {
				TcParams P;
				P.c = cb; P.c_row = c_row_bytes / 8; P.c_batch = kb_cstr / 8;
				P.n = (int)n; P.ntime = (int)ntime;
				int T = (int)div_up<long>(2 * n, 128);
				P.ntile = 0;
				for( int i=0; i<T; ++i ) P.ntile += i / 2 + 1;     // (row block, column-block pair) units
				P.alpha = (float)alpha; P.beta = (float)beta; P.conj_u = conj_u;
				static const bool use_cluster = getenv("BFB_LINALG_CLUSTER") != nullptr;
				static const int  nstage = getenv("BFB_LINALG_STAGES") ? atoi(getenv("BFB_LINALG_STAGES")) : 3;
				auto smem_for = [](int stages) {
					return (size_t)stages * TC_STAGE_BYTES + 64 * 65 * sizeof(float2) +
					       (2 * stages + 1) * sizeof(uint64_t) + 16 + 1024;   // + 1024: alignment of the window
				};
				if( use_cluster && T >= 2 ) {
					int TT = (T + 1) / 2;                           // 256-row super blocks
					size_t smem = smem_for(TC2_STAGES);
					BFB_CUDA(cudaFuncSetAttribute(corr_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
					                              (int)smem), BF_STATUS_INTERNAL_ERROR);
					dim3 grid2(2 * (TT * (TT + 1) / 2), (unsigned)kb);
					corr_tc2_kernel<<<grid2, TC_THREADS, smem, st>>>(tmap, P);
				} else {
					dim3 grid(P.ntile, (unsigned)kb);
#define BFB_TC_LAUNCH(S_) do { size_t smem = smem_for(S_); \
						BFB_CUDA(cudaFuncSetAttribute(corr_tc_kernel<S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
						                              (int)smem), BF_STATUS_INTERNAL_ERROR); \
						corr_tc_kernel<S_><<<grid, TC_THREADS, smem, st>>>(tmap, P); } while(0)
					switch( nstage ) {
					case 2:  BFB_TC_LAUNCH(2); break;
					case 4:  BFB_TC_LAUNCH(4); break;
					case 5:  BFB_TC_LAUNCH(5); break;
					case 8:  BFB_TC_LAUNCH(8); break;
					default: BFB_TC_LAUNCH(3); break;
					}
#undef BFB_TC_LAUNCH
				}
				count_launch();
				BFB_CUDA(cudaGetLastError(), BF_STATUS_INTERNAL_ERROR);
				continue;
			}
		}
		SimtParams S;
		S.u = ub; S.stride_t = stride_t; S.stride_i = stride_i; S.stride_b = kb_ustr;
		S.c = cb; S.c_row = c_row_bytes / 8; S.c_batch = kb_cstr / 8;
		S.n = (int)n; S.ntime = (int)ntime; S.nbatch = (int)kb;
		S.alpha = (float)alpha; S.beta = (float)beta; S.conj_u = conj_u;
		int T = (int)div_up<long>(n, 32);
		dim3 grid(T * (T + 1) / 2, (unsigned)kb);
		switch( utype ) {
		case BF_DTYPE_CI8:  corr_simt_kernel<int8_t, int><<<grid, 256, 0, st>>>(S); break;
		case BF_DTYPE_CI16: corr_simt_kernel<int16_t, float><<<grid, 256, 0, st>>>(S); break;
		case BF_DTYPE_CF32: corr_simt_kernel<float, float><<<grid, 256, 0, st>>>(S); break;
		default: BFB_FAIL(BF_STATUS_UNSUPPORTED_DTYPE);
		}
		count_launch();
		BFB_CUDA(cudaGetLastError(), BF_STATUS_INTERNAL_ERROR);
	}
	return BF_STATUS_SUCCESS;
}

} // namespace

// ---------------------------------------------------------------------------
// General product c = alpha a.b + beta c (reference: bfMatMul_ab,
// src/linalg.cu:479-637, cuBLAS / the beamformer kernel there).  A plain tiled
// SIMT GEMM with per-element loaders: any strides, conjugated views, and
// ci8 / ci16 / cf32 / f32 (float accumulation) or f64 / cf64 (double) inputs.
// It exists for API completeness (beamforming, reference test_linalg.py
// matmul_ab cases); the tensor-core work of this library is the correlator.
namespace bfb {

struct AbOperand { const char* p; long sm, sk, sb[BF_MAX_DIMS]; int kind, conj; };   // kind: BFdtype
struct AbParams {
	AbOperand a, b;                      // a[m][k], b[k][n] (strides in bytes: sm = row-like, sk = k)
	char* c; long c_m, c_n, c_b[BF_MAX_DIMS]; int c_kind;
	int  M, N, K, nb; long bshape[BF_MAX_DIMS];
	double alpha, beta;
};

template<typename R>
__device__ __forceinline__ void ab_load(const char* p, int kind, int conj, R& re, R& im) {
	switch( kind ) {
	case BF_DTYPE_CI8:  { char2  v = *(const char2*)p;   re = v.x; im = v.y; break; }
	case BF_DTYPE_CI16: { short2 v = *(const short2*)p;  re = v.x; im = v.y; break; }
	case BF_DTYPE_CF32: { float2 v = *(const float2*)p;  re = v.x; im = v.y; break; }
	case BF_DTYPE_CF64: { double2 v = *(const double2*)p; re = (R)v.x; im = (R)v.y; break; }
	case BF_DTYPE_F32:  re = *(const float*)p;  im = 0; break;
	case BF_DTYPE_F64:  re = (R)*(const double*)p; im = 0; break;
	case BF_DTYPE_I8:   re = *(const signed char*)p; im = 0; break;
	default:            re = 0; im = 0; break;
	}
	if( conj ) im = -im;
}

template<typename R>
__global__ void __launch_bounds__(256)
matmul_ab_kernel(AbParams P) {
	__shared__ R As[2][16][17], Bs[2][16][17];            // [re/im][m or k][k or n]
	long bidx = blockIdx.z, aoff = 0, boff = 0, coff = 0;
	for( int d=P.nb-1; d>=0; --d ) {
		long r = bidx % P.bshape[d]; bidx /= P.bshape[d];
		aoff += r * P.a.sb[d]; boff += r * P.b.sb[d]; coff += r * P.c_b[d];
	}
	const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
	const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
	R accr = 0, acci = 0;
	for( int k0=0; k0<P.K; k0+=16 ) {
		R re = 0, im = 0;
		if( m < P.M && k0 + tx < P.K )
			ab_load<R>(P.a.p + aoff + (long)m * P.a.sm + (long)(k0 + tx) * P.a.sk, P.a.kind, P.a.conj, re, im);
		As[0][ty][tx] = re; As[1][ty][tx] = im;
		re = 0; im = 0;
		if( n < P.N && k0 + ty < P.K )
			ab_load<R>(P.b.p + boff + (long)(k0 + ty) * P.b.sk + (long)n * P.b.sm, P.b.kind, P.b.conj, re, im);
		Bs[0][ty][tx] = re; Bs[1][ty][tx] = im;
		__syncthreads();
#pragma unroll
		for( int k=0; k<16; ++k ) {
			const R ar = As[0][ty][k], ai = As[1][ty][k], br = Bs[0][k][tx], bi = Bs[1][k][tx];
			accr += ar * br - ai * bi;
			acci += ar * bi + ai * br;
		}
		__syncthreads();
	}
	if( m >= P.M || n >= P.N ) return;
	char* cp = P.c + coff + (long)m * P.c_m + (long)n * P.c_n;
	R outr = (R)P.alpha * accr, outi = (R)P.alpha * acci;
	switch( P.c_kind ) {
	case BF_DTYPE_CF32: { float2* q = (float2*)cp; if( P.beta != 0 ) { outr += (R)P.beta * q->x; outi += (R)P.beta * q->y; }
	                      *q = make_float2((float)outr, (float)outi); break; }
	case BF_DTYPE_CF64: { double2* q = (double2*)cp; if( P.beta != 0 ) { outr += (R)P.beta * (R)q->x; outi += (R)P.beta * (R)q->y; }
	                      *q = make_double2((double)outr, (double)outi); break; }
	case BF_DTYPE_F32:  { float* q = (float*)cp; if( P.beta != 0 ) outr += (R)P.beta * *q; *q = (float)outr; break; }
	default:            { double* q = (double*)cp; if( P.beta != 0 ) outr += (R)P.beta * (R)*q; *q = (double)outr; break; }
	}
}

static bool ab_kind_ok(BFdtype t) {
	return t == BF_DTYPE_CI8 || t == BF_DTYPE_CI16 || t == BF_DTYPE_CF32 || t == BF_DTYPE_CF64 ||
	       t == BF_DTYPE_F32 || t == BF_DTYPE_F64 || t == BF_DTYPE_I8;
}

// The tensor-core form of the product, when the operands allow it (returns
// false to leave the call to the SIMT kernel): ci8 x ci8 -> cf32, b rows
// contiguous and 16-byte aligned (TMA), c rows contiguous, at most one batch
// dimension, sums that fit int32.
static bool matmul_ab_tc(BFlinalg_impl* h, AbParams const& P, long nbatch, BFstatus* status) {
	*status = BF_STATUS_SUCCESS;
	if( getenv("BFB_LINALG_SIMT") || !get_encode_fn() ) return false;
	if( P.a.kind != BF_DTYPE_CI8 || P.b.kind != BF_DTYPE_CI8 || P.c_kind != BF_DTYPE_CF32 ) return false;
	if( P.nb > 1 || P.c_n != 8 || P.c_m % 8 || P.b.sm != 2 || P.b.sk % 16 || ((uintptr_t)P.b.p % 16) ) return false;
	if( P.nb == 1 && (P.b.sb[0] % 16 || P.c_b[0] % 8 || P.b.sb[0] == 0) ) return false;
	if( (long)P.K * 2 * 127 * 127 >= (1L << 31) || P.K < 1 || P.M < 1 || P.N < 8 ) return false;
	cudaStream_t st = thread_stream();
	// a^T, k slow: [batch][K][pitch]
	const bool a_batched = P.nb == 1 && P.a.sb[0] != 0;
	const long na = a_batched ? nbatch : 1;
	const long pitch = round_up<long>(2L * P.M, 16);
	const size_t need = (size_t)na * P.K * pitch;
	if( h->at_size < need ) {
		cudaStreamSynchronize(st);
		if( h->at_buf ) cudaFree(h->at_buf);
		h->at_buf = nullptr; h->at_size = 0;
		if( cudaMalloc(&h->at_buf, need) != cudaSuccess ) { *status = BF_STATUS_MEM_ALLOC_FAILED; return true; }
		h->at_size = need;
	}
	{
		dim3 grid((unsigned)div_up<int>(P.K, 32), (unsigned)div_up<int>(P.M, 32), (unsigned)na);
		ab_transpose_a_kernel<<<grid, 256, 0, st>>>(P.a.p, P.a.sm, P.a.sk, a_batched ? P.a.sb[0] : 0,
		                                            (char*)h->at_buf, pitch, (long)P.K * pitch, P.M, P.K);
		count_launch();
	}
	CUtensorMap ta, tb;
	cuuint32_t box[3] = {128, TC_KT, 1}, estr[3] = {1, 1, 1};
	{
		cuuint64_t gdim[3] = {(cuuint64_t)(2 * P.M), (cuuint64_t)P.K, (cuuint64_t)na};
		cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)((long)P.K * pitch)};
		if( get_encode_fn()(&ta, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, h->at_buf, gdim, gstr, box, estr,
		                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
		                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ) return false;
	}
	{
		cuuint64_t gdim[3] = {(cuuint64_t)(2 * P.N), (cuuint64_t)P.K, (cuuint64_t)nbatch};
		cuuint64_t gstr[2] = {(cuuint64_t)P.b.sk, (cuuint64_t)(P.nb == 1 ? P.b.sb[0] : round_up<long>(P.b.sk * P.K, 16))};
		if( get_encode_fn()(&tb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)P.b.p, gdim, gstr, box, estr,
		                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
		                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ) return false;
	}
	AbTcParams Q;
	Q.c = (float2*)P.c; Q.c_row = P.c_m / 8; Q.c_batch = P.nb == 1 ? P.c_b[0] / 8 : 0;
	Q.M = P.M; Q.N = P.N; Q.K = P.K;
	Q.njj = div_up<int>(div_up<int>(P.N, 64), 2);
	Q.a_batched = a_batched ? 1 : 0;
	Q.alpha = (float)P.alpha; Q.beta = (float)P.beta;
	const bool ca = P.a.conj != 0, cb = P.b.conj != 0;
	Q.s_re = (ca != cb) ? 1.f : -1.f;
	Q.s_01 = cb ? -1.f : 1.f;
	Q.s_10 = ca ? -1.f : 1.f;
	constexpr int STAGES = 3;
	const size_t smem = (size_t)STAGES * TC_STAGE_BYTES + 64 * 65 * sizeof(float2) + (2 * STAGES + 1) * sizeof(uint64_t) + 16 + 1024;
	if( cudaFuncSetAttribute(ab_tc_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ) {
		*status = BF_STATUS_INTERNAL_ERROR; return true;
	}
	dim3 grid((unsigned)(div_up<int>(P.M, 64) * Q.njj), (unsigned)nbatch);
	ab_tc_kernel<STAGES><<<grid, TC_THREADS, smem, st>>>(ta, tb, Q);
	count_launch();
	if( cudaGetLastError() != cudaSuccess ) *status = BF_STATUS_INTERNAL_ERROR;
	return true;
}

static BFstatus matmul_ab(BFlinalg_impl* handle, double alpha, BFarray const* a, BFarray const* b, double beta, BFarray const* c) {
	BFB_ASSERT(space_on_device(a->space) && space_on_device(b->space), BF_STATUS_UNSUPPORTED_SPACE);
	int nd = c->ndim;
	BFB_ASSERT(a->ndim == nd && b->ndim == nd && nd >= 2 && nd <= BF_MAX_DIMS, BF_STATUS_INVALID_SHAPE);
	BFB_ASSERT(ab_kind_ok(a->dtype) && ab_kind_ok(b->dtype), BF_STATUS_UNSUPPORTED_DTYPE);
	BFB_ASSERT(c->dtype == BF_DTYPE_CF32 || c->dtype == BF_DTYPE_CF64 || c->dtype == BF_DTYPE_F32 ||
	           c->dtype == BF_DTYPE_F64, BF_STATUS_UNSUPPORTED_DTYPE);
	bool cplx_in = dtype_is_complex(a->dtype) || dtype_is_complex(b->dtype);
	BFB_ASSERT(!cplx_in || dtype_is_complex(c->dtype), BF_STATUS_INVALID_DTYPE);
	AbParams P;
	P.M = (int)c->shape[nd-2]; P.N = (int)c->shape[nd-1]; P.K = (int)a->shape[nd-1];
	BFB_ASSERT(a->shape[nd-2] == P.M && b->shape[nd-1] == P.N && b->shape[nd-2] == P.K, BF_STATUS_INVALID_SHAPE);
	P.a.p = (const char*)a->data; P.a.sm = a->strides[nd-2]; P.a.sk = a->strides[nd-1];
	P.a.kind = a->dtype; P.a.conj = a->conjugated ? 1 : 0;
	P.b.p = (const char*)b->data; P.b.sk = b->strides[nd-2]; P.b.sm = b->strides[nd-1];
	P.b.kind = b->dtype; P.b.conj = b->conjugated ? 1 : 0;
	P.c = (char*)c->data; P.c_m = c->strides[nd-2]; P.c_n = c->strides[nd-1]; P.c_kind = c->dtype;
	P.alpha = alpha; P.beta = beta;
	P.nb = 0;
	long nbatch = 1;
	for( int d=0; d<nd-2; ++d ) {
		BFB_ASSERT((a->shape[d] == c->shape[d] || a->shape[d] == 1) &&
		           (b->shape[d] == c->shape[d] || b->shape[d] == 1), BF_STATUS_INVALID_SHAPE);
		if( c->shape[d] == 1 ) continue;
		P.bshape[P.nb] = c->shape[d];
		P.a.sb[P.nb] = a->shape[d] == 1 ? 0 : a->strides[d];
		P.b.sb[P.nb] = b->shape[d] == 1 ? 0 : b->strides[d];
		P.c_b[P.nb] = c->strides[d];
		nbatch *= c->shape[d];
		++P.nb;
	}
	if( P.M == 0 || P.N == 0 || nbatch == 0 ) return BF_STATUS_SUCCESS;
	BFB_ASSERT(nbatch <= 65535, BF_STATUS_UNSUPPORTED_SHAPE);
	{
		BFstatus tcs = BF_STATUS_SUCCESS;
		if( matmul_ab_tc(handle, P, nbatch, &tcs) ) return tcs;
	}
	dim3 grid((unsigned)div_up<int>(P.N, 16), (unsigned)div_up<int>(P.M, 16), (unsigned)nbatch);
	bool dbl = a->dtype == BF_DTYPE_F64 || a->dtype == BF_DTYPE_CF64 || b->dtype == BF_DTYPE_F64 ||
	           b->dtype == BF_DTYPE_CF64;
	if( dbl ) matmul_ab_kernel<double><<<grid, 256, 0, thread_stream()>>>(P);
	else      matmul_ab_kernel<float ><<<grid, 256, 0, thread_stream()>>>(P);
	count_launch();
	BFB_CUDA(cudaGetLastError(), BF_STATUS_INTERNAL_ERROR);
	return BF_STATUS_SUCCESS;
}

} // namespace bfb

extern "C" {

BFstatus bfLinAlgCreate(BFlinalg* handle_ptr) {
	BFB_ASSERT(handle_ptr, BF_STATUS_INVALID_POINTER);
	*handle_ptr = nullptr;
	BFB_TRY(*handle_ptr = new BFlinalg_impl());
	return BF_STATUS_SUCCESS;
}

BFstatus bfLinAlgDestroy(BFlinalg handle) {
	BFB_ASSERT(handle, BF_STATUS_INVALID_HANDLE);
	delete handle;
	return BF_STATUS_SUCCESS;
}

BFstatus bfLinAlgMatMul(BFlinalg handle, double alpha, BFarray const* a, BFarray const* b,
                        double beta, BFarray const* c) {
	BFB_ASSERT(handle, BF_STATUS_INVALID_HANDLE);
	BFB_ASSERT(a || b, BF_STATUS_INVALID_ARGUMENT);
	BFB_ASSERT(c, BF_STATUS_INVALID_POINTER);
	BFB_ASSERT(space_on_device(c->space), BF_STATUS_UNSUPPORTED_SPACE);
	if( a && b ) { BFB_TRY(return matmul_ab(handle, alpha, a, b, beta, c)); }
	BFarray const* x = a ? a : b;
	BFB_ASSERT(space_on_device(x->space), BF_STATUS_UNSUPPORTED_SPACE);
	BFB_ASSERT(x->ndim == c->ndim && x->ndim >= 2 && x->ndim <= BF_MAX_DIMS, BF_STATUS_INVALID_SHAPE);
	BFB_ASSERT(c->dtype == BF_DTYPE_CF32, BF_STATUS_UNSUPPORTED_DTYPE);
	BFB_ASSERT(x->dtype == BF_DTYPE_CI8 || x->dtype == BF_DTYPE_CI16 || x->dtype == BF_DTYPE_CF32,
	           BF_STATUS_UNSUPPORTED_DTYPE);
	int nd = x->ndim;
	// b^H.b: x is [.., t, i];  a.a^H: x is [.., i, t] and u(t,i) = conj(a[i][t])
	long n      = a ? x->shape[nd-2] : x->shape[nd-1];
	long ntime  = a ? x->shape[nd-1] : x->shape[nd-2];
	long str_i  = a ? x->strides[nd-2] : x->strides[nd-1];
	long str_t  = a ? x->strides[nd-1] : x->strides[nd-2];
	int  conj_u = (a ? 1 : 0) ^ (x->conjugated ? 1 : 0);
	BFB_ASSERT(c->shape[nd-1] == n && c->shape[nd-2] == n, BF_STATUS_INVALID_SHAPE);
	BFB_ASSERT(c->strides[nd-1] == 8, BF_STATUS_UNSUPPORTED_STRIDE);
	long bshape[BF_MAX_DIMS], ustr[BF_MAX_DIMS], cstr[BF_MAX_DIMS];
	int nb = 0;
	for( int d=0; d<nd-2; ++d ) {
		BFB_ASSERT(x->shape[d] == c->shape[d] || x->shape[d] == 1, BF_STATUS_INVALID_SHAPE);
		if( c->shape[d] == 1 ) continue;
		bshape[nb] = c->shape[d];
		ustr[nb] = x->shape[d] == 1 ? 0 : x->strides[d];     // broadcast
		cstr[nb] = c->strides[d];
		++nb;
	}
	BFB_TRY(return correlate(x->dtype, x->data, str_t, str_i, conj_u, n, ntime, nb, bshape, ustr, cstr,
	                         c->data, c->strides[nd-2], alpha, beta, thread_stream()));
}

} // extern "C"
