// fdmt_chain.cuh -- the "chain" FDMT schedule for 1-byte inputs (v4).
//
// Two observations carry this schedule (the arithmetic contract is still the
// reference's, src/fdmt.cu:52-155, and the output is bit-identical):
//
// 1. EXACT INTEGERS.  With 8-bit input every value the transform ever forms is
//    an integer of magnitude <= 255*nchan.  For nchan <= 65536 that is below
//    2^24, so every fp32 add the reference performs is exact and the result
//    does not depend on the order or the width of the additions.  The lower
//    steps (sub-bands of <= 256 channels) therefore run on *packed pairs of
//    unsigned 16-bit accumulators* (one IADD = two samples, half the shared
//    memory and L2 bytes); signed input is biased by +128 per channel and the
//    bias (128*channels of the sub-band) is removed when a row is converted
//    to fp32.  Samples before t = 0 count as zero: x + 0 == x exactly, which
//    is the reference's "t >= delay" guard (fdmt.cu:133-139); the step-0 rows
//    with d > 0 (scaled running means, fdmt.cu:72-88, NaN for t < d) are not
//    integers -- a plan that uses one keeps the float schedule.
//
// 2. CHAINS.  out[d][t] = lo[dlo(d)][t] + hi[rest(d)][t - delay(d)]: the
//    low-frequency child is taken at the SAME time, only the high-frequency
//    child is shifted.  A warp that produces the rows of a sub-band in order
//    can therefore keep the low child's current row in registers, produce it
//    itself from *its* low child (again in registers) and so on down the left
//    spine of the merge tree; only high-frequency children are ever written
//    to shared memory.  One merge costs one (shifted) shared-memory read of
//    b instead of two reads and a write.
//
// A pass covers steps s0..s1.  A CTA owns a *program* = (sub-band of step s1,
// block of output delays) for T output samples; the host walks the tree and
// emits, per (program, level, warp), a list of ops  R[l] = (R[l-1] | smem a) +
// shift(smem b)  [-> smem | -> global].  Rows of one chain share a window
// [t0 - wmax, t0 + T - wmin); windows and all offsets are host-computed, the
// kernel has no delay logic.  Source rows are staged with TMA bulk copies
// (cp.async.bulk + mbarrier) from the previous pass's workspace, or converted
// from the 1-byte input.
#pragma once
#include "core.hpp"
#include "fdmt_plan.hpp"

#include <map>
#include <vector>
#include <algorithm>
#include <cstdint>

namespace bfb {

enum {
	CH_LEVEL_MASK = 7,
	CH_LOADA   = 1 << 3,     // a comes from shared memory (else registers of level-1)
	CH_NO_A    = 1 << 4,     // absent low-frequency parent
	CH_NO_B    = 1 << 5,     // absent high-frequency parent
	CH_STORE_S = 1 << 6,     // write the row to shared memory (chain head)
	CH_STORE_G = 1 << 7,     // write the row to the pass output
	CH_NVEC_SHIFT = 16,
	CH_MAXLEV  = 5,
	CH_LV      = 3,          // 16-byte vectors per lane per row
	CH_BYTE_WORDS = 6,       // 32-bit words per lane of the longest 1-byte source row (768 samples)
};
enum { CH_SRC_BYTES = 0, CH_SRC_SAME = 1 };
enum { CH_DST_SAME = 0, CH_DST_CVT = 1, CH_DST_FINAL = 2 };

struct ChainCfg {
	int D = 24;              // output delays per program
	int JR = 6;              // head rows per chain job
	int KD = 1;              // chain depth: levels a job keeps in registers (1: every row goes through shared memory)
	int nwarp = 8;
	int smem_cap = 110 * 1024;
	int tcap = 1 << 20;      // upper bound on T
};

struct ChainPass {
	int s0 = 0, s1 = 0, nlev = 0;
	int esize = 2;           // 2: packed u16 accumulators, 4: fp32
	bool chains = false;     // some op takes its a operand from registers
	int src_kind = CH_SRC_SAME, dst_kind = CH_DST_SAME;
	int T = 0, nprog = 0, nwarp = 0, slots = 0, src_slots = 0;
	int smem_elems = 0;      // data region (elements of esize bytes)
	int lookback = 0;        // largest backward reach of a source row (samples)
	int nrow_out = 0;        // rows of the pass output (compact index)
	long nops = 0;           // ops per time tile (all programs), for accounting
	std::vector<int4> ops;   // [prog][level-1][warp][slot]
	std::vector<int4> src;   // [prog][slot]: x row, y -smax, z smem offset, w length; w == 0 ends
	std::vector<int4> hdr;   // [prog]: x channels of the output band, y source rows, z staged bytes
	int4* d_ops = nullptr; int4* d_src = nullptr; int4* d_hdr = nullptr;
	int vs() const { return 16 / esize; }
	size_t smem_bytes() const {
		size_t b = ((size_t)nlev * nwarp * slots + src_slots + hdr_slots()) * sizeof(int4);
		b += 16;                                                     // mbarrier
		if( dst_kind == CH_DST_FINAL ) b += (size_t)nwarp * 32 * CH_LV * vs() * sizeof(float);
		return b + (size_t)smem_elems * esize;
	}
	static int hdr_slots() { return 1; }
};

namespace chain_detail {
struct Win { int lo, hi; };
struct SymOp { int level, row, flags, a_row, b_row, delay; };
struct Job {
	int level = 0; std::vector<int> rows; Win w = {0, 0}; std::vector<SymOp> ops; int warp = 0;
};
inline void grow(std::map<int, Win>& m, int row, int lo, int hi) {
	std::map<int, Win>::iterator it = m.find(row);
	if( it == m.end() ) { Win w = {lo, hi}; m[row] = w; }
	else { it->second.lo = std::min(it->second.lo, lo); it->second.hi = std::max(it->second.hi, hi); }
}
inline int fdiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline int floor_to(int a, int b) { return fdiv(a, b) * b; }
inline int ceil_to(int a, int b)  { return -floor_to(-a, b); }
} // namespace chain_detail

// Rows of every step that the final output depends on (the reference computes
// all rows; about a quarter of them are never read).
inline void fdmt_used_rows(FdmtPlan const& P, std::vector<std::vector<char> >* used_) {
	std::vector<std::vector<char> >& used = *used_;
	int S = P.nstep() - 1;
	used.assign(P.nstep(), std::vector<char>());
	for( int s=0; s<=S; ++s ) used[s].assign(P.nrow(s), 0);
	for( int r=0; r<P.nrow(S); ++r ) used[S][r] = 1;
	for( int s=S; s>=1; --s )
		for( int r=0; r<P.nrow(s); ++r ) if( used[s][r] ) {
			FdmtRow const& row = P.rows[s][r];
			if( row.src0 >= 0 ) used[s-1][row.src0] = 1;
			if( row.src1 >= 0 ) used[s-1][row.src1] = 1;
		}
}

// The integer schedule applies when no used step-0 row is a running mean
// (delay d > 0 inside its channel) and sums stay below 2^24.
inline bool fdmt_integer_safe(FdmtPlan const& P, std::vector<std::vector<char> > const& used) {
	if( (long)P.nchan * 255 >= (1L << 24) ) return false;
	for( size_t c=0; c<P.bands[0].size(); ++c )
		for( int d=1; d<P.bands[0][c].ndelay; ++d )
			if( used[0][P.bands[0][c].row0 + d] ) return false;
	return true;
}
// Last step whose sub-bands all fit 16-bit accumulators.
inline int fdmt_last_u16_step(FdmtPlan const& P) {
	int s16 = 0;
	for( int s=0; s<P.nstep(); ++s ) {
		bool ok = true;
		for( FdmtBand const& b : P.bands[s] ) if( b.nchan * 255 > 65535 ) ok = false;
		if( !ok ) break;
		s16 = s;
	}
	return s16;
}

// Builds the tables of the pass s0..s1.  `out_index[r]` maps a row of step s1
// to its row in the pass output (the compact workspace, or the delay itself
// for the final pass), `src_index[r]` a row of step s0-1 to its row in the
// source (workspace row, or input channel for the byte source).
inline bool build_chain_pass(FdmtPlan const& P, std::vector<std::vector<char> > const& used,
                             int s0, int s1, int esize, int src_kind, int dst_kind,
                             std::vector<int> const& src_index, std::vector<int> const& out_index,
                             ChainCfg const& cfg, ChainPass* cp) {
	using namespace chain_detail;
	if( s0 < 1 || s1 < s0 || s1 >= P.nstep() || s1 - s0 + 1 > CH_MAXLEV ) return false;
	const int nlev = s1 - s0 + 1;
	const int VS = 16 / esize, LS = CH_LV * VS, WLEN = 32 * LS;
	const int nwarp = cfg.nwarp;
	const int KD = std::max(1, std::min(cfg.KD, nlev));
	// row -> band per step of the pass
	std::vector<std::vector<int> > row_band(nlev + 1);
	for( int li=0; li<=nlev; ++li ) {
		int s = s0 - 1 + li;
		row_band[li].assign(P.nrow(s), 0);
		for( size_t b=0; b<P.bands[s].size(); ++b )
			for( int d=0; d<P.bands[s][b].ndelay; ++d ) row_band[li][P.bands[s][b].row0 + d] = (int)b;
	}
	// programs: blocks of used rows of each step-s1 band
	struct Prog { int band; std::vector<int> rows; };
	std::vector<Prog> progs;
	for( size_t b=0; b<P.bands[s1].size(); ++b ) {
		FdmtBand const& band = P.bands[s1][b];
		std::vector<int> rows;
		for( int d=0; d<band.ndelay; ++d ) if( used[s1][band.row0 + d] ) rows.push_back(band.row0 + d);
		if( rows.empty() ) continue;
		int nblk = std::max(1, div_up<int>((int)rows.size(), cfg.D));
		int bs   = div_up<int>((int)rows.size(), nblk);
		for( size_t i=0; i<rows.size(); i+=bs ) {
			Prog pg; pg.band = (int)b;
			pg.rows.assign(rows.begin() + i, rows.begin() + std::min(rows.size(), i + bs));
			progs.push_back(pg);
		}
	}
	if( progs.empty() ) return false;
	// heavy programs (many rows) first: their CTAs start first and the light
	// ones fill the tail of the launch
	std::stable_sort(progs.begin(), progs.end(), [&](Prog const& a, Prog const& b) {
		return P.bands[s1][a.band].ndelay * 64 + (int)a.rows.size() > P.bands[s1][b.band].ndelay * 64 + (int)b.rows.size();
	});
	struct Plan { std::vector<std::vector<Job> > jobs; std::map<int, Win> src; };
	std::vector<Plan> plans(progs.size());
	int max_spread = 0, lookback = 0;
	bool any_reg = false;
	for( size_t p=0; p<progs.size(); ++p ) {
		Plan& pl = plans[p];
		pl.jobs.assign(nlev + 1, std::vector<Job>());
		std::vector<std::map<int, Win> > req(nlev + 1);
		for( int r : progs[p].rows ) grow(req[nlev], r, 0, 0);
		for( int li=nlev; li>=1; --li ) {
			// group the required rows by band, cut each band's rows into jobs
			std::map<int, std::vector<int> > by_band;
			for( std::map<int, Win>::iterator it=req[li].begin(); it!=req[li].end(); ++it )
				by_band[row_band[li][it->first]].push_back(it->first);
			for( std::map<int, std::vector<int> >::iterator bt=by_band.begin(); bt!=by_band.end(); ++bt ) {
				std::vector<int>& rows = bt->second;          // ascending
				int jr   = (KD == 1) ? 1 : cfg.JR;            // without chains every row is its own job
				int njob = std::max(1, div_up<int>((int)rows.size(), jr));
				int per  = div_up<int>((int)rows.size(), njob);
				for( size_t i=0; i<rows.size(); i+=per ) {
					Job job; job.level = li;
					job.rows.assign(rows.begin() + i, rows.begin() + std::min(rows.size(), i + per));
					int lo = 1 << 30, hi = -(1 << 30);
					for( int r : job.rows ) { lo = std::min(lo, req[li][r].lo); hi = std::max(hi, req[li][r].hi); }
					job.w.lo = floor_to(lo, VS); job.w.hi = ceil_to(hi, VS);
					max_spread = std::max(max_spread, job.w.hi - job.w.lo);
					// generate the chain: rows of level li in order, low-frequency
					// descendants on demand down to KD levels, from shared memory below
					std::vector<int> cur(nlev + 1, -1);
					struct Gen {
						FdmtPlan const& P; int s0, KD; Job& job; std::vector<int>& cur;
						std::vector<std::map<int, Win> >& req; bool& any_reg;
						void row(int l, int r, int depth) {
							FdmtRow const& fr = P.rows[s0 - 1 + l][r];
							SymOp op; op.level = l; op.row = r; op.flags = 0;
							op.a_row = fr.src0; op.b_row = fr.src1; op.delay = fr.delay;
							if( fr.src0 < 0 ) op.flags |= CH_NO_A;
							else if( l == 1 || depth + 1 >= KD ) {
								op.flags |= CH_LOADA;
								grow(req[l-1], fr.src0, job.w.lo, job.w.hi);
							} else {
								any_reg = true;
								if( cur[l-1] != fr.src0 ) row(l - 1, fr.src0, depth + 1);
							}
							if( fr.src1 < 0 ) op.flags |= CH_NO_B;
							else grow(req[l-1], fr.src1, job.w.lo + fr.delay, job.w.hi + fr.delay);
							if( depth == 0 ) op.flags |= CH_STORE_S;     // refined to _G at encode time
							job.ops.push_back(op);
							cur[l] = r;
						}
					} gen = {P, s0, KD, job, cur, req, any_reg};
					for( int r : job.rows ) gen.row(li, r, 0);
					pl.jobs[li].push_back(job);
				}
			}
		}
		for( std::map<int, Win>::iterator it=req[0].begin(); it!=req[0].end(); ++it ) {
			Win w = { floor_to(it->second.lo, VS), ceil_to(it->second.hi, VS) };
			pl.src[it->first] = w;
			max_spread = std::max(max_spread, w.hi - w.lo);
			lookback = std::max(lookback, w.hi);
		}
	}
	int T = std::min(cfg.tcap, WLEN - max_spread) / 16 * 16;
	if( T < 64 ) return false;
	cp->s0 = s0; cp->s1 = s1; cp->nlev = nlev; cp->esize = esize;
	cp->src_kind = src_kind; cp->dst_kind = dst_kind; cp->chains = any_reg;
	cp->T = T; cp->nprog = (int)progs.size(); cp->nwarp = nwarp; cp->lookback = lookback;
	// warp assignment (longest job first onto the least loaded warp) and list lengths
	int slots = 1, src_slots = 1;
	long nops = 0;
	for( size_t p=0; p<progs.size(); ++p ) {
		src_slots = std::max(src_slots, (int)plans[p].src.size() + 1);
		for( int li=1; li<=nlev; ++li ) {
			std::vector<Job>& jobs = plans[p].jobs[li];
			std::vector<int> order(jobs.size());
			for( size_t j=0; j<jobs.size(); ++j ) order[j] = (int)j;
			std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return jobs[a].ops.size() > jobs[b].ops.size(); });
			std::vector<int> load(nwarp, 0);
			for( int j : order ) {
				int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
				jobs[j].warp = w; load[w] += (int)jobs[j].ops.size();
				nops += (long)jobs[j].ops.size();
			}
			slots = std::max(slots, *std::max_element(load.begin(), load.end()) + 1);
		}
	}
	cp->slots = slots; cp->src_slots = src_slots; cp->nops = nops;
	cp->ops.assign((size_t)cp->nprog * nlev * nwarp * slots, make_int4(0, 0, 0, 0));
	cp->src.assign((size_t)cp->nprog * src_slots, make_int4(0, 0, 0, 0));
	cp->hdr.assign((size_t)cp->nprog, make_int4(0, 0, 0, 0));
	// Shared-memory layout: level li lives in region li % nreg.  Rows of level li
	// are last read by chains headed at level li + KD and first overwritten in
	// phase li + nreg, so nreg = KD + 1 regions rotate safely.
	const int nreg = std::min(KD + 1, nlev);
	int smem_max = 0;
	for( size_t p=0; p<progs.size(); ++p ) {
		Plan& pl = plans[p];
		std::vector<std::map<int, int> > row_off(nlev + 1), row_job(nlev + 1);
		std::vector<int> level_size(nlev, 0);
		// pass 1: offsets inside each level
		{
			int off = 0;
			for( std::map<int, Win>::iterator it=pl.src.begin(); it!=pl.src.end(); ++it ) {
				int len = T + it->second.hi - it->second.lo;
				if( len > WLEN ) return false;
				row_off[0][it->first] = off;
				off += len + 2 * VS;
			}
			level_size[0] = off;
		}
		for( int li=1; li<nlev; ++li ) {
			int off = 0;
			for( size_t j=0; j<pl.jobs[li].size(); ++j ) {
				Job const& job = pl.jobs[li][j];
				int len = T + job.w.hi - job.w.lo;
				if( len > WLEN ) return false;
				for( int r : job.rows ) { row_off[li][r] = off; row_job[li][r] = (int)j; off += len + 2 * VS; }
			}
			level_size[li] = off;
		}
		std::vector<int> region_size(nreg, 0), region_base(nreg, 0);
		for( int li=0; li<nlev; ++li ) region_size[li % nreg] = std::max(region_size[li % nreg], level_size[li]);
		for( int r=1; r<nreg; ++r ) region_base[r] = region_base[r-1] + region_size[r-1];
		smem_max = std::max(smem_max, region_base[nreg-1] + region_size[nreg-1]);
		for( int li=0; li<nlev; ++li )
			for( std::map<int, int>::iterator it=row_off[li].begin(); it!=row_off[li].end(); ++it )
				it->second += region_base[li % nreg];
		// source table
		long staged = 0;
		int k = 0;
		for( std::map<int, Win>::iterator it=pl.src.begin(); it!=pl.src.end(); ++it, ++k ) {
			int len = T + it->second.hi - it->second.lo;
			if( it->first >= (int)src_index.size() || src_index[it->first] < 0 ) return false;
			cp->src[p * src_slots + k] = make_int4(src_index[it->first], -it->second.hi, row_off[0][it->first], len);
			staged += (long)len * esize;
		}
		cp->hdr[p] = make_int4(P.bands[s1][progs[p].band].nchan, (int)pl.src.size(), (int)staged, 0);
		// where a materialised row lives: offset, window
		struct Loc { int base, hi, len; };
		auto locate = [&](int li, int row) -> Loc {
			Loc L;
			L.base = row_off[li][row];
			if( li == 0 ) { Win const& sw = pl.src[row]; L.hi = sw.hi; L.len = T + sw.hi - sw.lo; }
			else { Job const& j = pl.jobs[li][row_job[li][row]]; L.hi = j.w.hi; L.len = T + j.w.hi - j.w.lo; }
			return L;
		};
		std::vector<int> fill((size_t)nlev * nwarp, 0);
		for( int li=1; li<=nlev; ++li )
			for( Job const& job : pl.jobs[li] ) {
				int len = T + job.w.hi - job.w.lo;
				for( SymOp const& so : job.ops ) {
					int4 op = make_int4(0, 0, 0, 0);
					int flags = so.flags;
					if( flags & CH_LOADA ) {
						Loc a = locate(so.level - 1, so.a_row);
						int ea = a.hi - job.w.hi;
						if( ea < 0 || ea % VS || ea + len > a.len ) return false;
						op.y = a.base + ea;
					}
					if( !(flags & CH_NO_B) ) {
						Loc b = locate(so.level - 1, so.b_row);
						int eb = b.hi - job.w.hi - so.delay;
						if( eb < 0 || eb + len > b.len ) return false;
						op.z = b.base + eb;
					}
					if( flags & CH_STORE_S ) {
						if( so.level == nlev ) {
							flags = (flags & ~CH_STORE_S) | CH_STORE_G;
							if( so.row >= (int)out_index.size() || out_index[so.row] < 0 ) return false;
							op.x = out_index[so.row];
						} else op.x = row_off[so.level][so.row];
					}
					op.w = so.level | flags | ((len / VS) << CH_NVEC_SHIFT);
					int& n = fill[(size_t)(job.level - 1) * nwarp + job.warp];
					cp->ops[(((size_t)p * nlev + (job.level - 1)) * nwarp + job.warp) * slots + n] = op;
					++n;
				}
			}
	}
	cp->smem_elems = smem_max + (32 * CH_LV + 4) * VS;     // slack: lanes past a row's end still load
	if( cp->smem_bytes() > (size_t)cfg.smem_cap ) return false;
	return true;
}

// ---------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------
struct ChainParams {
	const void* src; long sstride, sbatch, src_tb;   // source workspace (elements), time of column 0
	void*       dst; long dstride, dbatch, dst_tb;   // pass output
	const int4* ops; const int4* srcs; const int4* hdr;
	const void* raw; long rstride, rbatch;           // 1-byte input (elements)
	long ntime;                                      // samples of the gulp
	long t_begin;                                    // t0 of tile 0
	long ntile;
	int  T, nlev, slots, src_slots;
	int  is_signed;
};

namespace chain_dev {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template<int ESZ> __device__ __forceinline__ uint32_t ch_add(uint32_t a, uint32_t b) {
	if( ESZ == 2 ) return a + b;                       // two u16 lanes, no carry across (values bounded)
	return __float_as_uint(__fadd_rn(__uint_as_float(a), __uint_as_float(b)));
}

// b's 12 words for this lane: 16 loaded words, sub-vector shift `sub` elements.
template<int ESZ, int WO>
__device__ __forceinline__ void ch_shift(const uint32_t (&bw)[16], int hbits, uint32_t (&t)[12]) {
#pragma unroll
	for( int k=0; k<12; ++k )
		t[k] = (ESZ == 2) ? __funnelshift_r(bw[k + WO], bw[k + WO + 1], hbits) : bw[k + WO];
}
} // namespace chain_dev

// A finished row of the pass's top level -> pass output (same element type),
// or, for an fp32 pass that ends the plan, the diagonal store of fdmt.cu:141-147.
template<int ESZ, int DSTK>
__device__ __forceinline__ void ch_store_out(const uint32_t (&o)[12], const int4& op, int nvec,
                                             const ChainParams& P, long t0, float* scratch, int lane, int warp) {
	constexpr int VS = 16 / ESZ, LS = CH_LV * VS;
	if( DSTK == CH_DST_FINAL ) {
		float* sc = scratch + (size_t)warp * 32 * LS;
		__syncwarp();
#pragma unroll
		for( int j=0; j<CH_LV; ++j )
			*(uint4*)(sc + LS * lane + 4 * j) = make_uint4(o[4*j], o[4*j+1], o[4*j+2], o[4*j+3]);
		__syncwarp();
		const long d = op.x;
		float* g = (float*)P.dst + (long)blockIdx.z * P.dbatch + d * P.dstride - d + t0;
#pragma unroll 4
		for( int i=lane; i<P.T; i+=32 ) {
			const long t = t0 + i;
			if( t >= d && t < P.ntime ) g[i] = sc[i];
		}
	} else {
		unsigned char* g = (unsigned char*)P.dst +
			((long)blockIdx.z * P.dbatch + (long)op.x * P.dstride + (t0 - P.dst_tb)) * ESZ;
		uint4* d = (uint4*)g + CH_LV * lane;
#pragma unroll
		for( int j=0; j<CH_LV; ++j )
			if( CH_LV * lane + j < nvec ) d[j] = make_uint4(o[4*j], o[4*j+1], o[4*j+2], o[4*j+3]);
	}
}
// Top level of a 16-bit pass with fp32 output: halves added in 32 bits, bias
// removed, converted (exact), stored to the fp32 workspace or diagonally.
template<int DSTK>
__device__ __forceinline__ void ch_store_wide(const uint32_t (&av)[12], const uint32_t (&bt)[12], const int4& op,
                                              int nvec, int bias, const ChainParams& P, long t0,
                                              float* scratch, int lane, int warp) {
	constexpr int LS = CH_LV * 8;
	float f[2 * 12];
#pragma unroll
	for( int k=0; k<12; ++k ) {
		int lo = (int)(av[k] & 0xFFFFu) + (int)(bt[k] & 0xFFFFu) - bias;
		int hi = (int)(av[k] >> 16)     + (int)(bt[k] >> 16)     - bias;
		f[2*k] = (float)lo; f[2*k+1] = (float)hi;
	}
	if( DSTK == CH_DST_CVT ) {
		float* g = (float*)P.dst + (long)blockIdx.z * P.dbatch + (long)op.x * P.dstride + (t0 - P.dst_tb) + LS * lane;
#pragma unroll
		for( int j=0; j<2*CH_LV; ++j )
			if( 2 * (CH_LV * lane) + j < 2 * nvec )
				*(float4*)(g + 4 * j) = make_float4(f[4*j], f[4*j+1], f[4*j+2], f[4*j+3]);
	} else {
		float* sc = scratch + (size_t)warp * 32 * LS;
		__syncwarp();
#pragma unroll
		for( int j=0; j<2*CH_LV; ++j ) *(float4*)(sc + LS * lane + 4 * j) = make_float4(f[4*j], f[4*j+1], f[4*j+2], f[4*j+3]);
		__syncwarp();
		const long d = op.x;
		float* g = (float*)P.dst + (long)blockIdx.z * P.dbatch + d * P.dstride - d + t0;
#pragma unroll 4
		for( int i=lane; i<P.T; i+=32 ) {
			const long t = t0 + i;
			if( t >= d && t < P.ntime ) g[i] = sc[i];
		}
	}
}

// Fetch / produce one level's registers with compile-time indices.
template<int L, int NLMAX>
__device__ __forceinline__ void ch_get(const uint32_t (&R)[NLMAX + 1][12], uint32_t (&av)[12]) {
	if constexpr( L >= 1 && L <= NLMAX ) {
#pragma unroll
		for( int k=0; k<12; ++k ) av[k] = R[L][k];
	}
}
// R[L] = (use_av ? av : R[L-1]) + bt, then the optional store of a chain head.
template<int ESZ, int DSTK, int L, int NLMAX>
__device__ __forceinline__ void ch_level(uint32_t (&R)[NLMAX + 1][12], const uint32_t (&av)[12],
                                         const uint32_t (&bt)[12], bool use_av, const int4& op,
                                         unsigned char* dbase, const ChainParams& P, long t0,
                                         float* scratch, int lane, int warp) {
	using namespace chain_dev;
	if constexpr( L >= 1 && L <= NLMAX ) {
		constexpr int VS = 16 / ESZ, LS = CH_LV * VS;
		if( use_av || L == 1 ) {
#pragma unroll
			for( int k=0; k<12; ++k ) R[L][k] = ch_add<ESZ>(av[k], bt[k]);
		} else {
#pragma unroll
			for( int k=0; k<12; ++k ) R[L][k] = ch_add<ESZ>(R[L-1][k], bt[k]);
		}
		const int nvec = op.w >> CH_NVEC_SHIFT;
		if( op.w & CH_STORE_S ) {
			uint4* d = (uint4*)(dbase + (size_t)op.x * ESZ) + CH_LV * lane;
#pragma unroll
			for( int j=0; j<CH_LV; ++j )
				if( CH_LV * lane + j < nvec ) d[j] = make_uint4(R[L][4*j], R[L][4*j+1], R[L][4*j+2], R[L][4*j+3]);
		} else if( op.w & CH_STORE_G ) {
			if( DSTK == CH_DST_FINAL ) {           // fp32 pass, last plan step: diagonal store
				float* sc = scratch + (size_t)warp * 32 * LS;
				__syncwarp();
#pragma unroll
				for( int j=0; j<CH_LV; ++j )
					*(uint4*)(sc + LS * lane + 4 * j) = make_uint4(R[L][4*j], R[L][4*j+1], R[L][4*j+2], R[L][4*j+3]);
				__syncwarp();
				const long d = op.x;
				float* g = (float*)P.dst + (long)blockIdx.z * P.dbatch + d * P.dstride - d + t0;
#pragma unroll 4
				for( int i=lane; i<P.T; i+=32 ) {
					const long t = t0 + i;
					if( t >= d && t < P.ntime ) g[i] = sc[i];
				}
			} else {
				unsigned char* g = (unsigned char*)P.dst +
					((long)blockIdx.z * P.dbatch + (long)op.x * P.dstride + (t0 - P.dst_tb)) * ESZ;
				uint4* d = (uint4*)g + CH_LV * lane;
#pragma unroll
				for( int j=0; j<CH_LV; ++j )
					if( CH_LV * lane + j < nvec ) d[j] = make_uint4(R[L][4*j], R[L][4*j+1], R[L][4*j+2], R[L][4*j+3]);
			}
		}
	}
}

// One op of a pass without register chains: row = a + shift(b), both from
// shared memory, to shared memory or to the pass output.
template<int ESZ, int DSTK>
__device__ __forceinline__ void ch_row_op(const int4& op, unsigned char* dbase, const ChainParams& P, long t0,
                                          int bias, float* scratch, int lane, int warp) {
	using namespace chain_dev;
	constexpr int VS = 16 / ESZ;
	const int nvec = op.w >> CH_NVEC_SHIFT;
	uint32_t av[12], bw[16];
	if( op.w & (CH_NO_A | CH_NO_B) ) {
		// absent parent (odd band counts): zeros stand in
		if( op.w & CH_NO_A ) {
#pragma unroll
			for( int k=0; k<12; ++k ) av[k] = 0u;
		} else {
			const uint4* a = (const uint4*)(dbase + (size_t)op.y * ESZ) + CH_LV * lane;
#pragma unroll
			for( int j=0; j<CH_LV; ++j ) { uint4 v = a[j]; av[4*j] = v.x; av[4*j+1] = v.y; av[4*j+2] = v.z; av[4*j+3] = v.w; }
		}
		if( op.w & CH_NO_B ) {
#pragma unroll
			for( int k=0; k<16; ++k ) bw[k] = 0u;
		} else {
			const int sub = op.z & (VS - 1);
			const uint4* b = (const uint4*)(dbase + (size_t)(op.z - sub) * ESZ) + CH_LV * lane;
#pragma unroll
			for( int j=0; j<CH_LV+1; ++j ) { uint4 v = b[j]; bw[4*j] = v.x; bw[4*j+1] = v.y; bw[4*j+2] = v.z; bw[4*j+3] = v.w; }
		}
	} else {
		const uint4* a = (const uint4*)(dbase + (size_t)op.y * ESZ) + CH_LV * lane;
		const int sub = op.z & (VS - 1);
		const uint4* b = (const uint4*)(dbase + (size_t)(op.z - sub) * ESZ) + CH_LV * lane;
#pragma unroll
		for( int j=0; j<CH_LV; ++j ) { uint4 v = a[j]; av[4*j] = v.x; av[4*j+1] = v.y; av[4*j+2] = v.z; av[4*j+3] = v.w; }
#pragma unroll
		for( int j=0; j<CH_LV+1; ++j ) { uint4 v = b[j]; bw[4*j] = v.x; bw[4*j+1] = v.y; bw[4*j+2] = v.z; bw[4*j+3] = v.w; }
	}
	const int sub = (op.w & CH_NO_B) ? 0 : (op.z & (VS - 1));
	const int wo = (ESZ == 2) ? (sub >> 1) : sub;
	const int hbits = (ESZ == 2) ? (sub & 1) * 16 : 0;
	if( (ESZ == 2) && (DSTK != CH_DST_SAME) && (op.w & CH_STORE_G) ) {
		// top level of a 16-bit pass whose output is fp32: the sum may exceed
		// 16 bits, add the halves in 32 bits and drop the bias
		uint32_t bt[12];
		switch( wo ) {
		case 0:  ch_shift<ESZ, 0>(bw, hbits, bt); break;
		case 1:  ch_shift<ESZ, 1>(bw, hbits, bt); break;
		case 2:  ch_shift<ESZ, 2>(bw, hbits, bt); break;
		default: ch_shift<ESZ, 3>(bw, hbits, bt); break;
		}
		ch_store_wide<DSTK>(av, bt, op, nvec, bias, P, t0, scratch, lane, warp);
		return;
	}
	switch( wo ) {
#define BFB_CH_CASE(W_) \
		_Pragma("unroll") for( int k=0; k<12; ++k ) \
			av[k] = ch_add<ESZ>(av[k], (ESZ == 2) ? __funnelshift_r(bw[k + W_], bw[k + W_ + 1], hbits) : bw[k + W_]);
	case 0:  BFB_CH_CASE(0) break;
	case 1:  BFB_CH_CASE(1) break;
	case 2:  BFB_CH_CASE(2) break;
	default: BFB_CH_CASE(3) break;
#undef BFB_CH_CASE
	}
	if( op.w & CH_STORE_S ) {
		uint4* d = (uint4*)(dbase + (size_t)op.x * ESZ) + CH_LV * lane;
#pragma unroll
		for( int j=0; j<CH_LV; ++j )
			if( CH_LV * lane + j < nvec ) d[j] = make_uint4(av[4*j], av[4*j+1], av[4*j+2], av[4*j+3]);
	} else {
		ch_store_out<ESZ, DSTK>(av, op, nvec, P, t0, scratch, lane, warp);
	}
}

template<int ESZ, int SRCK, int DSTK, int NLMAX>     // NLMAX == 0: no op keeps a row in registers
__global__ void __launch_bounds__(256, NLMAX == 0 ? 3 : 2)
fdmt_chain_kernel(const __grid_constant__ ChainParams P) {
	using namespace chain_dev;
	constexpr int VS = 16 / ESZ;           // elements per 16-byte vector
	constexpr int LS = CH_LV * VS;         // samples per lane
	extern __shared__ __align__(16) unsigned char ch_smem[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
	int4* sops = (int4*)ch_smem;
	const int nop = P.nlev * nwarp * P.slots;
	int4* ssrc = sops + nop;
	int4* shdr = ssrc + P.src_slots;
	uint64_t* mbar = (uint64_t*)(shdr + 1);
	float* scratch = (float*)(mbar + 2);
	unsigned char* dbase = (unsigned char*)(scratch + (DSTK == CH_DST_FINAL ? nwarp * 32 * LS : 0));
	{
		const int4* g = P.ops + (size_t)blockIdx.y * nop;
		for( int i=threadIdx.x; i<nop; i+=blockDim.x ) sops[i] = __ldg(g + i);
		const int4* gs = P.srcs + (size_t)blockIdx.y * P.src_slots;
		for( int i=threadIdx.x; i<P.src_slots; i+=blockDim.x ) ssrc[i] = __ldg(gs + i);
		if( threadIdx.x == 0 ) {
			shdr[0] = __ldg(P.hdr + blockIdx.y);
			mbar_init(mbar, 1);
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
	}
	__syncthreads();
	const int4 hdr = shdr[0];
	const int bias = P.is_signed ? 128 * hdr.x : 0;
	const unsigned char* src = (const unsigned char*)P.src + (long)blockIdx.z * P.sbatch * ESZ;
	const unsigned char* rin = (const unsigned char*)P.raw + (long)blockIdx.z * P.rbatch;
	uint32_t parity = 0;

	for( long tile=blockIdx.x; tile<P.ntile; tile+=gridDim.x ) {
		const long t0 = P.t_begin + tile * P.T;
		// ---- stage the source rows
		if( SRCK == CH_SRC_BYTES ) {
			// Two rows per warp at a time, all of their global loads issued before
			// the first is consumed (the loads come from HBM: ~1 us each).
			uint16_t* data = (uint16_t*)dbase;
			const uint32_t flip = P.is_signed ? 0x80808080u : 0u;
			for( int k0=warp; k0<hdr.y; k0+=2*nwarp ) {
				uint32_t w[2][CH_BYTE_WORDS + 1];
				int4 e[2]; bool fast[2];
#pragma unroll
				for( int q=0; q<2; ++q ) {
					const int k = k0 + q * nwarp;
					e[q] = (k < hdr.y) ? ssrc[k] : make_int4(0, 0, 0, 0);
					const long ts = t0 + e[q].y;
					const unsigned char* g = rin + (long)e[q].x * P.rstride + ts;
					fast[q] = e[q].w > 0 && ts >= 4 && ts + e[q].w + 4 <= P.ntime;
					const uint32_t* ga = (const uint32_t*)(g - ((uintptr_t)g & 3));
					const int nword = e[q].w >> 2;
#pragma unroll
					for( int i=0; i<=CH_BYTE_WORDS; ++i ) {
						const int j = lane + 32 * i;
						w[q][i] = (fast[q] && j <= nword) ? __ldg(ga + j) : 0u;
					}
				}
#pragma unroll
				for( int q=0; q<2; ++q ) {
					if( e[q].w == 0 ) continue;
					const long ts = t0 + e[q].y;
					const unsigned char* g = rin + (long)e[q].x * P.rstride + ts;
					uint16_t* srow = data + e[q].z;
					const int nword = e[q].w >> 2;
					if( fast[q] ) {
						const unsigned mis = (unsigned)((uintptr_t)g & 3);
#pragma unroll
						for( int i=0; i<CH_BYTE_WORDS; ++i ) {
							const int j = lane + 32 * i;
							uint32_t hi = __shfl_down_sync(0xffffffffu, w[q][i], 1);
							const uint32_t nx = __shfl_sync(0xffffffffu, w[q][i + 1], 0);
							if( lane == 31 ) hi = nx;
							if( j < nword ) {
								const uint32_t v = __funnelshift_r(w[q][i], hi, mis * 8) ^ flip;
								*(uint2*)(srow + 4 * j) = make_uint2(__byte_perm(v, 0, 0x4140), __byte_perm(v, 0, 0x4342));
							}
						}
					} else {
						for( int j=lane; j<nword; j+=32 ) {
							uint32_t v = 0;
#pragma unroll
							for( int b=0; b<4; ++b ) {
								const long t = ts + 4 * j + b;
								const uint32_t x = (t >= 0 && t < P.ntime) ? (uint32_t)g[4 * j + b] : 0u;
								v |= x << (8 * b);
							}
							v ^= flip;
							*(uint2*)(srow + 4 * j) = make_uint2(__byte_perm(v, 0, 0x4140), __byte_perm(v, 0, 0x4342));
						}
					}
				}
			}
		} else {
			if( warp == 0 ) {
				if( lane == 0 ) { fence_proxy_async(); mbar_expect_tx(mbar, (uint32_t)hdr.z); }
				__syncwarp();
				for( int k=lane; k<hdr.y; k+=32 ) {
					const int4 e = ssrc[k];
					const unsigned char* g = src + ((long)e.x * P.sstride + (t0 + e.y - P.src_tb)) * ESZ;
					bulk_g2s(dbase + (size_t)e.z * ESZ, g, (uint32_t)e.w * ESZ, mbar);
				}
			}
			while( !mbar_try_wait(mbar, parity) ) { }
			parity ^= 1;
		}
		__syncthreads();

		// ---- merge levels
		if constexpr( NLMAX == 0 ) {
			for( int lev=1; lev<=P.nlev; ++lev ) {
				const int4* list = sops + ((size_t)(lev - 1) * nwarp + warp) * P.slots;
				int4 nxt = list[0];
				for( int m=0; m<P.slots; ++m ) {
					const int4 op = nxt;
					if( op.w == 0 ) break;
					nxt = list[m + 1];                     // the last slot of a list is always a terminator
					ch_row_op<ESZ, DSTK>(op, dbase, P, t0, bias, scratch, lane, warp);
				}
				__syncthreads();
			}
		} else {
		uint32_t R[NLMAX + 1][12];
		for( int lev=1; lev<=P.nlev; ++lev ) {
			const int4* list = sops + ((size_t)(lev - 1) * nwarp + warp) * P.slots;
			for( int m=0; m<P.slots; ++m ) {
				const int4 op = list[m];
				if( op.w == 0 ) break;
				const int l = op.w & CH_LEVEL_MASK;
				uint32_t av[12], bt[12];
				bool use_av = false;
				if( op.w & CH_LOADA ) {
					const uint4* a = (const uint4*)(dbase + (size_t)op.y * ESZ) + CH_LV * lane;
#pragma unroll
					for( int j=0; j<CH_LV; ++j ) { uint4 v = a[j]; av[4*j] = v.x; av[4*j+1] = v.y; av[4*j+2] = v.z; av[4*j+3] = v.w; }
					use_av = true;
				} else if( op.w & CH_NO_A ) {
#pragma unroll
					for( int k=0; k<12; ++k ) av[k] = 0u;
					use_av = true;
				}
				if( !(op.w & CH_NO_B) ) {
					const int sub = op.z & (VS - 1);
					const uint4* b = (const uint4*)(dbase + (size_t)(op.z - sub) * ESZ) + CH_LV * lane;
					uint32_t bw[16];
#pragma unroll
					for( int j=0; j<CH_LV+1; ++j ) { uint4 v = b[j]; bw[4*j] = v.x; bw[4*j+1] = v.y; bw[4*j+2] = v.z; bw[4*j+3] = v.w; }
					const int wo = (ESZ == 2) ? (sub >> 1) : sub;
					const int hbits = (ESZ == 2) ? (sub & 1) * 16 : 0;
					switch( wo ) {
					case 0:  ch_shift<ESZ, 0>(bw, hbits, bt); break;
					case 1:  ch_shift<ESZ, 1>(bw, hbits, bt); break;
					case 2:  ch_shift<ESZ, 2>(bw, hbits, bt); break;
					default: ch_shift<ESZ, 3>(bw, hbits, bt); break;
					}
				} else {
#pragma unroll
					for( int k=0; k<12; ++k ) bt[k] = 0u;
				}
				if( (op.w & CH_STORE_G) && (ESZ == 2) && (DSTK != CH_DST_SAME) ) {
					// top level of a 16-bit pass whose output is fp32: the sum may
					// exceed 16 bits, add the halves in 32 bits and drop the bias
					if( !use_av ) {
						switch( l ) {
						case 2: ch_get<1, NLMAX>(R, av); break;
						case 3: ch_get<2, NLMAX>(R, av); break;
						case 4: ch_get<3, NLMAX>(R, av); break;
						case 5: ch_get<4, NLMAX>(R, av); break;
						default: break;
						}
					}
					const int nvec = op.w >> CH_NVEC_SHIFT;
					float f[2 * 12];
#pragma unroll
					for( int k=0; k<12; ++k ) {
						int lo = (int)(av[k] & 0xFFFFu) + (int)(bt[k] & 0xFFFFu) - bias;
						int hi = (int)(av[k] >> 16)     + (int)(bt[k] >> 16)     - bias;
						f[2*k] = (float)lo; f[2*k+1] = (float)hi;
					}
					if( DSTK == CH_DST_CVT ) {
						float* g = (float*)P.dst + (long)blockIdx.z * P.dbatch + (long)op.x * P.dstride + (t0 - P.dst_tb) + LS * lane;
#pragma unroll
						for( int j=0; j<2*CH_LV; ++j )
							if( 2 * (CH_LV * lane) + j < 2 * nvec )
								*(float4*)(g + 4 * j) = make_float4(f[4*j], f[4*j+1], f[4*j+2], f[4*j+3]);
					} else {
						float* sc = scratch + (size_t)warp * 32 * LS;
						__syncwarp();
#pragma unroll
						for( int j=0; j<2*CH_LV; ++j ) *(float4*)(sc + LS * lane + 4 * j) = make_float4(f[4*j], f[4*j+1], f[4*j+2], f[4*j+3]);
						__syncwarp();
						const long d = op.x;
						float* g = (float*)P.dst + (long)blockIdx.z * P.dbatch + d * P.dstride - d + t0;
#pragma unroll 4
						for( int i=lane; i<P.T; i+=32 ) {
							const long t = t0 + i;
							if( t >= d && t < P.ntime ) g[i] = sc[i];
						}
					}
					continue;
				}
				switch( l ) {
				case 1:  ch_level<ESZ, DSTK, 1, NLMAX>(R, av, bt, use_av, op, dbase, P, t0, scratch, lane, warp); break;
				case 2:  ch_level<ESZ, DSTK, 2, NLMAX>(R, av, bt, use_av, op, dbase, P, t0, scratch, lane, warp); break;
				case 3:  ch_level<ESZ, DSTK, 3, NLMAX>(R, av, bt, use_av, op, dbase, P, t0, scratch, lane, warp); break;
				case 4:  ch_level<ESZ, DSTK, 4, NLMAX>(R, av, bt, use_av, op, dbase, P, t0, scratch, lane, warp); break;
				default: ch_level<ESZ, DSTK, 5, NLMAX>(R, av, bt, use_av, op, dbase, P, t0, scratch, lane, warp); break;
				}
			}
			__syncthreads();
		}
		}   // chains
	}
}

} // namespace bfb
