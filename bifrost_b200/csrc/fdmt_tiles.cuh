// fdmt_tiles.cuh -- fused FDMT merge passes over (band, delay-block, time-tile)
// tiles held in shared memory.
//
// The reference runs every merge step of src/fdmt.cu:95-155 as its own launch
// (one HBM round trip of the whole state per step, fdmt.cu:698-716).  Here a
// pass covers several consecutive steps s0..s1: a CTA owns the output rows
// [d_lo, d_hi) of one step-s1 sub-band for T output samples, walks the merge
// tree downwards on the host to find exactly which rows of every lower step it
// needs and over which time window, stages the step-(s0-1) rows once, and runs
// the steps between two shared-memory regions.
//
// Windows are *per row*: a row needed with time shifts [smin, smax] relative to
// the tile's output times is held for times [t0 - smax, t0 + T - smin), so the
// halo is the spread of the delays inside the delay block, not their size.
// With that, one merge is  dst[w] = a[w + ea] + b[w + eb]  for the row's whole
// window with host-computed constant offsets -- no delay test in the kernel:
// samples before t = 0 are staged as -0.0f, and x + (-0.0f) == x bit for bit,
// which reproduces the reference's "t >= delay" guard (fdmt.cu:133-139).
// smax is rounded up (smin down) to multiples of 4 so that dst and a are
// 16-byte aligned and only b carries a sub-vector shift (-delay & 3).
// Arithmetic is the reference's single fp32 add per output, so results are
// bit-identical.
#pragma once
#include "core.hpp"
#include "fdmt_plan.hpp"

#include <map>
#include <vector>
#include <algorithm>

namespace bfb {

enum { TILE_NO_A = 1, TILE_NO_B = 2, TILE_SLOW = 4 };
enum { TILE_VPL = 3 };                       // float4 per lane per row
enum { TILE_LMAX = 128 * TILE_VPL };         // longest row window (floats)

// Work item (int4).  Stage phase:  x global row, y time offset of sample 0
// relative to t0 (<= 0, multiple of 4), z shared offset, w nvec.
// Merge phase: x dst shared offset (last phase: global row), y a offset,
// z b offset (any alignment), w nvec | flags << 16.   nvec == 0 ends a list.
// Raw pass (s0 == 1): the stage phase copies input channels (x input row,
// y time offset, z byte offset, w bytes) and the step-1 items read them through
// the step-0 running means (fdmt.cu:52-92): y/z are byte offsets,
// w = nvec | dd0 << 8 | dd1 << 12 | flags << 16, and aux holds
// (smax of the row, delay, exact dd0, exact dd1) for the exact edge path.
struct TilePass {
	int s0 = 0, s1 = 0, T = 0, nprog = 0, nphase = 0, nwarp = 0, slots = 0;
	int smem_floats = 0;                     // float regions
	bool raw = false;                        // s0 == 1: stage the 1-byte input itself
	int raw_bytes = 0;                       // shared bytes in front of the float regions
	int edge_margin = 0;                     // tiles with t0 < edge_margin reach t < 0
	std::vector<int4> items;                 // [prog][phase][warp][slot]
	std::vector<int4> aux;                   // raw pass, step-1 items: [prog][warp][slot]
	int4* d_items = nullptr;
	int4* d_aux = nullptr;
};

struct TileParams {
	const float* src; long sstride, sbatch;  // step s0-1 rows
	float* dst;       long dstride, dbatch;  // step s1 rows, or the output array
	const int4* items;
	const int4* aux;                         // raw pass only
	const void* raw; long rstride, rbatch;   // raw pass: input array (elements)
	int  raw_bytes, edge_margin;
	long ntime;
	long src_limit;                          // readable floats per src row
	long dst_limit;                          // writable floats per dst row (non-final)
	int  T, nphase, slots;
	long ntile;                              // time tiles in the gulp
	int  tiles_per_cta;                      // consecutive tiles walked by one CTA
};

namespace tile_detail {
struct Range { int smin, smax, off; };
typedef std::map<int, Range> RowMap;

inline void merge_range(RowMap& m, int row, int lo, int hi) {
	RowMap::iterator it = m.find(row);
	if( it == m.end() ) { Range r = {lo, hi, 0}; m[row] = r; }
	else { it->second.smin = std::min(it->second.smin, lo); it->second.smax = std::max(it->second.smax, hi); }
}

// need[0] = step s0-1 rows ... need[nlev-1] = step s1 rows of the block
inline void walk(FdmtPlan const& P, int s0, int s1, int row_lo, int row_hi, std::vector<RowMap>* need_) {
	std::vector<RowMap>& need = *need_;
	int nlev = s1 - s0 + 2;
	need.assign(nlev, RowMap());
	for( int r=row_lo; r<row_hi; ++r ) merge_range(need[nlev-1], r, 0, 0);
	for( int li=nlev-1; li>=0; --li ) {
		int s = s0 - 1 + li;
		for( RowMap::iterator it=need[li].begin(); it!=need[li].end(); ++it ) {
			Range& g = it->second;
			g.smin = g.smin & ~3;
			g.smax = (g.smax + 3) & ~3;
			if( li == 0 ) continue;
			FdmtRow const& row = P.rows[s][it->first];
			if( row.src0 >= 0 ) merge_range(need[li-1], row.src0, g.smin, g.smax);
			if( row.src1 >= 0 ) merge_range(need[li-1], row.src1, g.smin + row.delay, g.smax + row.delay);
		}
	}
}
} // namespace tile_detail

// Builds the item tables of the pass s0..s1 with delay blocks of about D rows.
// Returns false when the pass cannot be tiled (window too wide).
inline bool build_tile_pass(FdmtPlan const& P, int s0, int s1, int D, int nwarp, TilePass* tp, bool raw=false) {
	using namespace tile_detail;
	if( s0 < 1 || s1 < s0 || s1 >= P.nstep() ) return false;
	if( raw && s0 != 1 ) return false;
	// step-0 row -> (channel, delay)
	std::vector<int> row_chan, row_dd;
	if( raw ) {
		row_chan.resize(P.nrow(0)); row_dd.resize(P.nrow(0));
		for( size_t c=0; c<P.bands[0].size(); ++c )
			for( int d=0; d<P.bands[0][c].ndelay; ++d ) {
				row_chan[P.bands[0][c].row0 + d] = (int)c;
				row_dd  [P.bands[0][c].row0 + d] = d;
			}
	}
	struct Prog { int row_lo, row_hi; };
	std::vector<Prog> progs;
	for( size_t b=0; b<P.bands[s1].size(); ++b ) {
		FdmtBand const& band = P.bands[s1][b];
		int nblk = std::max(1, div_up<int>(band.ndelay, D));
		int bs   = div_up<int>(band.ndelay, nblk);
		for( int d=0; d<band.ndelay; d+=bs ) {
			Prog pg = {band.row0 + d, band.row0 + std::min(band.ndelay, d + bs)};
			progs.push_back(pg);
		}
	}
	int nlev = s1 - s0 + 2;
	std::vector<std::vector<RowMap> > needs(progs.size());
	int max_spread = 0;
	size_t max_rows = 0;
	int max_shift = 0;
	std::vector<RowMap> chans(progs.size());     // raw pass: channel windows
	for( size_t p=0; p<progs.size(); ++p ) {
		walk(P, s0, s1, progs[p].row_lo, progs[p].row_hi, &needs[p]);
		for( int li=(raw ? 1 : 0); li<nlev; ++li ) {
			max_rows = std::max(max_rows, needs[p][li].size());
			for( RowMap::iterator it=needs[p][li].begin(); it!=needs[p][li].end(); ++it ) {
				max_spread = std::max(max_spread, it->second.smax - it->second.smin);
				max_shift  = std::max(max_shift, it->second.smax);
			}
		}
		if( raw ) {
			for( RowMap::iterator it=needs[p][0].begin(); it!=needs[p][0].end(); ++it )
				merge_range(chans[p], row_chan[it->first], it->second.smin, it->second.smax + row_dd[it->first]);
			for( RowMap::iterator it=chans[p].begin(); it!=chans[p].end(); ++it ) {
				it->second.smin &= ~3;
				it->second.smax = (it->second.smax + 3) & ~3;
				max_shift = std::max(max_shift, it->second.smax);
			}
			max_rows = std::max(max_rows, chans[p].size());
		}
	}
	int T = (TILE_LMAX - max_spread) & ~3;
	if( T < 64 ) return false;
	tp->s0 = s0; tp->s1 = s1; tp->T = T; tp->nprog = (int)progs.size();
	tp->nphase = nlev; tp->nwarp = nwarp; tp->raw = raw; tp->edge_margin = max_shift + 8;
	tp->slots = div_up<int>((int)max_rows, nwarp) + 1;          // +1: terminator
	tp->items.assign((size_t)tp->nprog * tp->nphase * nwarp * tp->slots, make_int4(0, 0, 0, 0));
	if( raw ) tp->aux.assign((size_t)tp->nprog * nwarp * tp->slots, make_int4(0, 0, 0, 0));
	int smem_max = 0, raw_max = 0;
	for( size_t p=0; p<progs.size(); ++p ) {
		std::vector<RowMap>& need = needs[p];
		// shared-memory offsets: even levels in region 0, odd levels in region 1
		int region[2] = {0, 0};
		if( raw ) {
			// channel rows (bytes): 16 guard bytes, rows back to back, 16 slack bytes
			int off = 16;
			for( RowMap::iterator it=chans[p].begin(); it!=chans[p].end(); ++it ) {
				it->second.off = off;
				if( T + it->second.smax - it->second.smin > 508 ) return false;   // tile_bytes_load
				off += T + it->second.smax - it->second.smin;
			}
			raw_max = std::max(raw_max, (off + 32 + 15) & ~15);
		}
		for( int li=(raw ? 1 : 0); li<nlev-1; ++li ) {          // the last level goes to HBM
			int off = 0;
			for( RowMap::iterator it=need[li].begin(); it!=need[li].end(); ++it ) {
				it->second.off = off;
				off += T + it->second.smax - it->second.smin;
			}
			region[li & 1] = std::max(region[li & 1], off + 16); // slack: whole-lane loads + b's look-ahead
		}
		for( int li=1; li<nlev-1; li+=2 )
			for( RowMap::iterator it=need[li].begin(); it!=need[li].end(); ++it ) it->second.off += region[0];
		smem_max = std::max(smem_max, region[0] + region[1]);
		for( int li=0; li<nlev; ++li ) {
			int4* base = &tp->items[(((size_t)p * tp->nphase + li) * nwarp) * tp->slots];
			int k = 0;
			if( raw && li == 0 ) {
				for( RowMap::iterator it=chans[p].begin(); it!=chans[p].end(); ++it, ++k ) {
					Range const& g = it->second;
					int c_in = P.reverse_band ? P.nchan - 1 - it->first : it->first;
					base[(size_t)(k % nwarp) * tp->slots + (k / nwarp)] =
						make_int4(c_in, -g.smax, g.off, T + g.smax - g.smin);
				}
				continue;
			}
			if( raw && li == 1 ) {
				int4* abase = &tp->aux[((size_t)p * nwarp) * tp->slots];
				for( RowMap::iterator it=need[li].begin(); it!=need[li].end(); ++it, ++k ) {
					Range const& g = it->second;
					size_t si = (size_t)(k % nwarp) * tp->slots + (k / nwarp);
					int nvec = (T + g.smax - g.smin) / 4;
					FdmtRow const& row = P.rows[1][it->first];
					int flags = 0, a_off = 0, b_off = 0, dd0 = 0, dd1 = 0;
					if( row.src0 >= 0 ) {
						Range const& q = chans[p][row_chan[row.src0]];
						dd0 = row_dd[row.src0];
						a_off = q.off + (q.smax - g.smax);
					} else flags |= TILE_NO_A;
					if( row.src1 >= 0 ) {
						Range const& q = chans[p][row_chan[row.src1]];
						dd1 = row_dd[row.src1];
						b_off = q.off + (q.smax - g.smax - row.delay);
					} else flags |= TILE_NO_B;
					if( dd0 > 3 || dd1 > 3 || flags ) flags |= TILE_SLOW;
					int x = (nlev == 2) ? it->first : g.off;
					base[si]  = make_int4(x, a_off, b_off, nvec | (std::min(dd0, 3) << 8) | (std::min(dd1, 3) << 12) | (flags << 16));
					abase[si] = make_int4(g.smax, row.delay, dd0, dd1);
				}
				continue;
			}
			for( RowMap::iterator it=need[li].begin(); it!=need[li].end(); ++it, ++k ) {
				Range const& g = it->second;
				int4* slot = base + (size_t)(k % nwarp) * tp->slots + (k / nwarp);
				int nvec = (T + g.smax - g.smin) / 4;
				if( li == 0 ) {
					*slot = make_int4(it->first, -g.smax, g.off, nvec);
					continue;
				}
				int s = s0 - 1 + li;
				FdmtRow const& row = P.rows[s][it->first];
				int flags = 0, a_off = 0, b_off = 0;
				if( row.src0 >= 0 ) {
					Range const& q = need[li-1][row.src0];
					a_off = q.off + (q.smax - g.smax);
				} else flags |= TILE_NO_A;
				if( row.src1 >= 0 ) {
					Range const& q = need[li-1][row.src1];
					b_off = q.off + (q.smax - g.smax - row.delay);
				} else flags |= TILE_NO_B;
				int x = (li == nlev-1) ? it->first : g.off;
				*slot = make_int4(x, a_off, b_off, nvec | (flags << 16));
			}
		}
	}
	tp->smem_floats = smem_max;
	tp->raw_bytes = raw_max;
	return true;
}

// ---------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------
template<int SH>
__device__ __forceinline__ float4 tile_shifted(float4 lo, float4 hi) {
	if( SH == 0 ) return lo;
	if( SH == 1 ) return make_float4(lo.y, lo.z, lo.w, hi.x);
	if( SH == 2 ) return make_float4(lo.z, lo.w, hi.x, hi.y);
	return make_float4(lo.w, hi.x, hi.y, hi.z);
}

__device__ __forceinline__ float4 tile_add(float4 a, float4 b) {
	return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}

// One row, shared -> shared.  Lane l owns vectors 3l .. 3l+2 (a 48-byte lane
// stride keeps 16-byte accesses conflict-free) and carries b's upper vector
// to the next one: VPL+1 loads of b instead of 2*VPL.
template<int SH>
__device__ __forceinline__ void tile_row_smem(float4* __restrict__ d, const float4* __restrict__ a,
                                              const float4* __restrict__ b, int nv) {
	// Lanes at the end of a row own fewer than VPL vectors: they still load
	// and add all of them (the regions carry slack for that) and only the
	// stores are predicated, so the warp never runs a second code path.
	if( nv <= 0 ) return;
	float4 va[TILE_VPL], vb[TILE_VPL + 1];
#pragma unroll
	for( int j=0; j<TILE_VPL; ++j ) { va[j] = a[j]; vb[j] = b[j]; }
	vb[TILE_VPL] = SH ? b[TILE_VPL] : vb[0];
#pragma unroll
	for( int j=0; j<TILE_VPL; ++j ) {
		float4 o = tile_add(va[j], tile_shifted<SH>(vb[j], vb[j + 1]));
		if( j < nv ) d[j] = o;
	}
}

// One row, shared -> global row (pass output), lane-consecutive vectors.
template<int SH>
__device__ __forceinline__ void tile_row_out(float* __restrict__ g, long room_vec, const float4* __restrict__ a,
                                             const float4* __restrict__ b, int nvec, int lane) {
#pragma unroll
	for( int j=0; j<TILE_VPL; ++j ) {
		int v = lane + 32 * j;
		if( v < nvec && v < room_vec ) {
			float4 lo = b[v], hi = SH ? b[v + 1] : lo;
			*(float4*)(g + 4 * v) = tile_add(a[v], tile_shifted<SH>(lo, hi));
		}
	}
}

// ---- raw pass: step 1 straight from the staged 1-byte input ------------------
// Staging of one input-channel row: `nbyte` bytes (multiple of 4, at most 508)
// from global `g` (any alignment) to a 4-aligned shared row, zero-filling
// bytes outside [valid_lo, valid_hi) (byte offsets relative to g).  Split in
// a load half (aligned 32-bit words into registers, nothing consumes them, so
// the loads of several rows and of the next tile stay in flight) and a store
// half (funnel shift with the neighbour lane's word, then shared stores).
// Rows that touch the ends of the array take the exact byte path at store time.
__device__ __forceinline__ bool tile_bytes_interior(int nbyte, long valid_lo, long valid_hi) {
	return valid_lo <= -4 && valid_hi >= (long)nbyte + 8;
}
__device__ __forceinline__ void tile_bytes_load(const unsigned char* __restrict__ g, int nbyte,
                                                long valid_lo, long valid_hi, int lane, uint32_t (&word)[4]) {
	const int nword = nbyte >> 2;
	const uint32_t* ga = (const uint32_t*)(g - ((uintptr_t)g & 3));
	const bool interior = tile_bytes_interior(nbyte, valid_lo, valid_hi);
#pragma unroll
	for( int k=0; k<4; ++k ) {
		const int j = k * 32 + lane;
		word[k] = (interior && j <= nword) ? __ldg(ga + j) : 0u;
	}
}
__device__ __forceinline__ void tile_bytes_store(const unsigned char* __restrict__ g,
                                                 unsigned char* __restrict__ srow, int nbyte,
                                                 long valid_lo, long valid_hi, int lane,
                                                 const uint32_t (&word)[4]) {
	uint32_t* sw = (uint32_t*)srow;
	const int nword = nbyte >> 2;
	const unsigned mis = (unsigned)((uintptr_t)g & 3);
	if( tile_bytes_interior(nbyte, valid_lo, valid_hi) ) {
#pragma unroll
		for( int k=0; k<4; ++k ) {
			const int j = k * 32 + lane;
			uint32_t hi = __shfl_down_sync(0xffffffffu, word[k], 1);
			uint32_t nx = __shfl_sync(0xffffffffu, word[(k + 1) & 3], 0);
			if( lane == 31 ) hi = nx;
			if( j < nword ) sw[j] = __funnelshift_r(word[k], hi, mis * 8);
		}
		return;
	}
	for( int j=lane; j<nword; j+=32 ) {
		const long b0 = (long)j * 4;
		uint32_t w = 0;
#pragma unroll
		for( int q=0; q<4; ++q ) {
			long bb = b0 + q;
			uint32_t v = (bb >= valid_lo && bb < valid_hi) ? (uint32_t)g[bb] : 0u;
			w |= v << (8 * q);
		}
		sw[j] = w;
	}
}

template<bool SIGNED>
__device__ __forceinline__ float tile_byte(uint32_t w, int k) {
	return SIGNED ? (float)(signed char)(w >> (8 * k)) : (float)(unsigned char)(w >> (8 * k));
}

// Step-0 running means (delay DD <= 3) of the 12 samples that follow the
// first word: w[0] holds samples -4..-1, w[1..3] samples 0..11.  Same fp32
// operation order as fdmt.cu:72-88 (sum newest to oldest, one multiply).
template<int DD, bool SIGNED>
__device__ __forceinline__ void tile_state0(const uint32_t (&w)[4], float (&f)[12]) {
	float x[16];
#pragma unroll
	for( int k=4-DD; k<16; ++k ) x[k] = tile_byte<SIGNED>(w[k >> 2], k & 3);
	const float scale = __fdiv_rn(1.f, (float)(DD + 1));
#pragma unroll
	for( int j=0; j<12; ++j ) {
		float acc = x[4 + j];
#pragma unroll
		for( int k=1; k<=DD; ++k ) acc = __fadd_rn(acc, x[4 + j - k]);
		f[j] = DD ? __fmul_rn(acc, scale) : acc;
	}
}

// One step-1 row for this lane's 12 samples.  `a` (4-aligned) and `b` (any
// alignment) point at the lane's first sample inside the staged channels.
template<int D0, int D1, bool SIGNED>
__device__ __forceinline__ void tile_row_raw(float (&o)[12], const unsigned char* __restrict__ a,
                                             const unsigned char* __restrict__ b) {
	uint32_t wa[4], wb[4], t[5];
	const uint32_t* pa = (const uint32_t*)(a - 4);
#pragma unroll
	for( int k=0; k<4; ++k ) wa[k] = pa[k];
	const unsigned sh = (unsigned)((uintptr_t)b & 3);
	const uint32_t* pb = (const uint32_t*)(b - sh - 4);
#pragma unroll
	for( int k=0; k<5; ++k ) t[k] = pb[k];
#pragma unroll
	for( int k=0; k<4; ++k ) wb[k] = __funnelshift_r(t[k], t[k + 1], sh * 8);
	float fa[12], fb[12];
	tile_state0<D0, SIGNED>(wa, fa);
	tile_state0<D1, SIGNED>(wb, fb);
#pragma unroll
	for( int j=0; j<12; ++j ) o[j] = __fadd_rn(fa[j], fb[j]);
}

template<bool SIGNED>
__device__ __forceinline__ void tile_row_raw_dispatch(int dd, float (&o)[12], const unsigned char* a,
                                                      const unsigned char* b) {
	switch( dd ) {             // dd0 | dd1 << 4, warp-uniform
#define BFB_RAW_CASE(D0_, D1_) case (D0_) | ((D1_) << 4): tile_row_raw<D0_, D1_, SIGNED>(o, a, b); break;
	BFB_RAW_CASE(0,0) BFB_RAW_CASE(1,0) BFB_RAW_CASE(2,0) BFB_RAW_CASE(3,0)
	BFB_RAW_CASE(0,1) BFB_RAW_CASE(1,1) BFB_RAW_CASE(2,1) BFB_RAW_CASE(3,1)
	BFB_RAW_CASE(0,2) BFB_RAW_CASE(1,2) BFB_RAW_CASE(2,2) BFB_RAW_CASE(3,2)
	BFB_RAW_CASE(0,3) BFB_RAW_CASE(1,3) BFB_RAW_CASE(2,3)
	default: tile_row_raw<3, 3, SIGNED>(o, a, b); break;
#undef BFB_RAW_CASE
	}
}

// Exact step-0 value at absolute time t (fdmt.cu:72-88): NaN for t < dd, and
// this file's -0.0f convention before t = 0.  x points at the sample of time t.
template<bool SIGNED>
__device__ __forceinline__ float tile_state0_exact(const unsigned char* x, int dd, long t) {
	if( t < 0 )  return -0.f;
	if( t < dd ) return CUDART_NAN_F;
	float acc = 0.f;
	for( int k=0; k<=dd; ++k )
		acc = __fadd_rn(acc, SIGNED ? (float)(signed char)x[-k] : (float)x[-k]);
	return __fmul_rn(acc, __fdiv_rn(1.f, (float)(dd + 1)));
}

// RAW: 0 = the source is the float state of step s0-1; 1 / 2 = the source is
// the signed / unsigned 1-byte input array itself (s0 == 1).
template<bool FINAL, int RAW>
__global__ void __launch_bounds__(256, 3)
fdmt_tile_kernel(const __grid_constant__ TileParams P) {
	extern __shared__ __align__(16) float tsmem[];
	const int  lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
	// This program's work items are copied to shared memory once per CTA (a
	// dependent global load per item would dominate the row loops); the CTA
	// then walks `tiles_per_cta` consecutive time tiles.
	int4* sitems = (int4*)tsmem;
	const int nitem = P.nphase * nwarp * P.slots;
	{
		const int4* gitems = P.items + (size_t)blockIdx.y * nitem;
		for( int i=threadIdx.x; i<nitem; i+=blockDim.x ) sitems[i] = __ldg(gitems + i);
	}
	unsigned char* rbuf = (unsigned char*)(sitems + nitem);  // staged input channels (raw pass)
	float* tbuf = (float*)(rbuf + (RAW ? P.raw_bytes : 0));
	const float* src = P.src + (long)blockIdx.z * P.sbatch;
	float*       dst = P.dst + (long)blockIdx.z * P.dbatch;
	const unsigned char* rin = (const unsigned char*)P.raw + (long)blockIdx.z * P.rbatch;
	const int4* const stage_items = sitems + warp * P.slots;
	__syncthreads();

	// Staging keeps the loads of TILE_NB rows per warp in flight before any of
	// them is consumed (prefetching the next tile's rows across the merge
	// phases was measured to add nothing on top of that).
	constexpr int TILE_NB = RAW ? 5 : 3;
	auto raw_row = [&](const int4& it, long t0n, const unsigned char*& g, long& lo, long& hi) {
		const long wstart = t0n + it.y;
		g = rin + (long)it.x * P.rstride + wstart;        // may point before the row
		lo = wstart < 0 ? -wstart : 0; hi = P.ntime - wstart;
		if( wstart >= 4 ) lo = -4;                        // the aligned word in front belongs to the array
	};
	auto load_stage = [&](long t0n, int m, uint32_t (&rw)[4], float4 (&rv)[TILE_VPL]) {
		const int4 it = stage_items[m];
		if( RAW ) {
			const unsigned char* g; long lo, hi;
			raw_row(it, t0n, g, lo, hi);
			tile_bytes_load(g, it.w, lo, hi, lane, rw);
		} else {
			// step-(s0-1) rows: -0.0f before t = 0, zeros past the row
			const long tb = t0n + it.y;
			const float* g = src + (long)it.x * P.sstride + tb;
#pragma unroll
			for( int j=0; j<TILE_VPL; ++j ) {
				const int v = lane + 32 * j;
				const long t = tb + 4 * v;
				float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
				if( v < it.w ) {
					if( t < 0 )                     val = make_float4(-0.f, -0.f, -0.f, -0.f);
					else if( t + 4 <= P.src_limit ) val = __ldg((const float4*)(g + 4 * v));
				}
				rv[j] = val;
			}
		}
	};
	auto store_stage = [&](long t0n, int m, const uint32_t (&rw)[4], const float4 (&rv)[TILE_VPL]) {
		const int4 it = stage_items[m];
		if( RAW ) {
			const unsigned char* g; long lo, hi;
			raw_row(it, t0n, g, lo, hi);
			tile_bytes_store(g, rbuf + it.z, it.w, lo, hi, lane, rw);
		} else {
			float4* sdst = (float4*)(tbuf + it.z);
#pragma unroll
			for( int j=0; j<TILE_VPL; ++j ) {
				const int v = lane + 32 * j;
				if( v < it.w ) sdst[v] = rv[j];
			}
		}
	};
	int nstage = 0;                                   // this warp's stage rows
	while( nstage < P.slots && stage_items[nstage].w != 0 ) ++nstage;

	const long tile_begin = (long)blockIdx.x * P.tiles_per_cta;
	const long tile_end   = min(P.ntile, tile_begin + P.tiles_per_cta);
	for( long tile=tile_begin; tile<tile_end; ++tile ) {
	const long t0 = tile * P.T;
	const int4* items = stage_items;
	for( int m0=0; m0<nstage; m0+=TILE_NB ) {
		uint32_t rw[RAW ? TILE_NB : 1][4];
		float4   rv[RAW ? 1 : TILE_NB][TILE_VPL];
#pragma unroll
		for( int q=0; q<TILE_NB; ++q ) if( m0 + q < nstage ) load_stage(t0, m0 + q, rw[RAW ? q : 0], rv[RAW ? 0 : q]);
#pragma unroll
		for( int q=0; q<TILE_NB; ++q ) if( m0 + q < nstage ) store_stage(t0, m0 + q, rw[RAW ? q : 0], rv[RAW ? 0 : q]);
	}
	__syncthreads();

	for( int phase=1; phase<P.nphase; ++phase ) {
		items += nwarp * P.slots;
		const bool last = (phase == P.nphase - 1);
		for( int m=0; m<P.slots; ++m ) {
			const int4 it = items[m];
			const int nvec = it.w & 0xFFFF;
			if( nvec == 0 ) break;
			const int flags = it.w >> 16;
			if( RAW && phase == 1 ) {
				const int nv12 = it.w & 0xFF;                   // vectors in the row
				const unsigned char* a = rbuf + it.y;
				const unsigned char* b = rbuf + it.z;
				const bool exact = (flags != 0) || (t0 < P.edge_margin);
				if( !exact ) {
					const int nv = nv12 - TILE_VPL * lane;
					if( nv > 0 ) {
						float o[12];
						tile_row_raw_dispatch<RAW == 1>((it.w >> 8) & 0xFF, o, a + 12 * lane, b + 12 * lane);
						if( !last ) {
							float4* d = (float4*)(tbuf + it.x) + TILE_VPL * lane;
#pragma unroll
							for( int j=0; j<TILE_VPL; ++j )
								if( j < nv ) d[j] = make_float4(o[4*j], o[4*j+1], o[4*j+2], o[4*j+3]);
						} else {
#pragma unroll
							for( int j=0; j<12; ++j ) {
								long t = t0 + 12 * lane + j;
								if( j < 4 * nv ) {
									if( !FINAL ) { if( t < P.dst_limit ) dst[(long)it.x * P.dstride + t] = o[j]; }
									else if( t >= it.x && t < P.ntime ) dst[(long)it.x * P.dstride + (t - it.x)] = o[j];
								}
							}
						}
					}
				} else {
					// exact edge semantics (t < delay, NaN rows, absent parents, dd > 3)
					const int4 ax = P.aux[((size_t)blockIdx.y * nwarp + warp) * P.slots + m];
					const long tw = t0 - ax.x;                      // time of the row's sample 0
					for( int w=lane; w<4*nv12; w+=32 ) {
						const long t = tw + w;
						float va = (flags & TILE_NO_A) ? 0.f : tile_state0_exact<RAW == 1>(a + w, ax.z, t);
						float val = va;
						if( !(flags & TILE_NO_B) ) {
							float vb = tile_state0_exact<RAW == 1>(b + w, ax.w, t - ax.y);
							val = __fadd_rn(va, vb);
						}
						if( (flags & TILE_NO_A) && t < 0 ) val = -0.f;
						if( !last ) tbuf[it.x + w] = val;
						else if( !FINAL ) { if( t >= 0 && t < P.dst_limit ) dst[(long)it.x * P.dstride + t] = val; }
						else if( t >= it.x && t < P.ntime ) dst[(long)it.x * P.dstride + (t - it.x)] = val;
					}
				}
				continue;
			}
			if( flags ) {
				// absent parent (odd band counts): one sample per lane
				const float* a = tbuf + it.y;
				const float* b = tbuf + it.z;
				for( int w=lane; w<4*nvec; w+=32 ) {
					float val = (flags & TILE_NO_A) ? 0.f : a[w];
					if( !(flags & TILE_NO_B) ) val = __fadd_rn(val, b[w]);
					long t = t0 + w;
					if( !last ) tbuf[it.x + w] = val;
					else if( !FINAL ) { if( t < P.dst_limit ) dst[(long)it.x * P.dstride + t] = val; }
					else if( t >= it.x && t < P.ntime ) dst[(long)it.x * P.dstride + (t - it.x)] = val;
				}
				continue;
			}
			const int sh = it.z & 3;
			if( !last ) {
				float4* d = (float4*)(tbuf + it.x) + TILE_VPL * lane;
				const float4* a = (const float4*)(tbuf + it.y) + TILE_VPL * lane;
				const float4* b = (const float4*)(tbuf + (it.z & ~3)) + TILE_VPL * lane;
				const int nv = nvec - TILE_VPL * lane;
				switch( sh ) {
				case 0:  tile_row_smem<0>(d, a, b, nv); break;
				case 1:  tile_row_smem<1>(d, a, b, nv); break;
				case 2:  tile_row_smem<2>(d, a, b, nv); break;
				default: tile_row_smem<3>(d, a, b, nv); break;
				}
			} else if( !FINAL ) {
				float* g = dst + (long)it.x * P.dstride + t0;
				const long room = (P.dst_limit - t0) >> 2;
				const float4* a = (const float4*)(tbuf + it.y);
				const float4* b = (const float4*)(tbuf + (it.z & ~3));
				switch( sh ) {
				case 0:  tile_row_out<0>(g, room, a, b, nvec, lane); break;
				case 1:  tile_row_out<1>(g, room, a, b, nvec, lane); break;
				case 2:  tile_row_out<2>(g, room, a, b, nvec, lane); break;
				default: tile_row_out<3>(g, room, a, b, nvec, lane); break;
				}
			} else {
				// last plan step: row d is stored shifted left by d (fdmt.cu:141-147)
				const float* a = tbuf + it.y;
				const float* b = tbuf + it.z;
				const long d = it.x;
				float* g = dst + d * P.dstride - d + t0;
				const int  w_hi = (int)min((long)(4 * nvec), P.ntime - t0);
#pragma unroll 4
				for( int w=lane; w<w_hi; w+=32 )
					if( t0 + w >= d ) g[w] = __fadd_rn(a[w], b[w]);
			}
		}
		__syncthreads();
	}
	}   // tiles
}

} // namespace bfb
