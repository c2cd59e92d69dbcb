// fdmt_plan.hpp -- host-side FDMT plan: the sub-band merge tree and, for every
// step, the (source row, source row, delay) triple of each output row.
//
// Restates the algorithm of the reference's BFfdmt_impl::init
// (src/fdmt.cu:338-530) with its own data structures.  The tables it produces
// must be identical to the reference's, because they *define* which samples
// are summed; tests/test_fdmt_plan.py pins them against the numpy oracle.
#pragma once

#include <cmath>
#include <complex>
#include <limits>
#include <vector>

namespace bfb {

struct FdmtBand {          // one sub-band at one step
	int chan0;             // first channel
	int nchan;             // channels covered
	int row0;              // first state row at this step
	int ndelay;            // rows (delays 0..ndelay-1)
	int parent0, parent1;  // bands of the previous step (-1 = absent)
};

struct FdmtRow {           // one output row of a merge step
	int src0;              // row of the low-frequency parent (-1 = none)
	int src1;              // row of the high-frequency parent (-1 = none)
	int delay;             // time shift applied to src1
};

struct FdmtPlan {
	int    nchan = 0, max_delay = 0;
	double f0 = 0, df = 0, exponent = 0;
	bool   reverse_band = false;
	std::vector<std::vector<FdmtBand> > bands;   // [step][band]
	std::vector<std::vector<FdmtRow> >  rows;    // [step][row]; rows[0] empty
	int nrow_max = 0;

	int nstep() const { return (int)bands.size(); }
	int nrow(int step) const {
		FdmtBand const& b = bands[step].back();
		return b.row0 + b.ndelay;
	}

	double chan_freq(int c) const { return f0 + df * c; }

	// (flo^g - fhi^g) / (fmin^g - fmax^g), evaluated in complex arithmetic so
	// that negative frequencies work (ref: src/fdmt.cu:301-319).
	double rel_delay(double flo, double fhi, double fmin, double fmax) const {
		typedef std::complex<double> C;
		C numer = std::pow(C(flo),  exponent) - std::pow(C(fhi),  exponent);
		C denom = std::pow(C(fmin), exponent) - std::pow(C(fmax), exponent);
		double eps = std::numeric_limits<double>::epsilon();
		if( std::norm(denom) < eps*eps ) return 0;
		return (numer / denom).real();
	}
	double rel_delay_band(double flo, double fhi) const {
		return rel_delay(flo, fhi, chan_freq(0), chan_freq(nchan-1));
	}
	int band_ndelay(double flo, double width) const {
		double frac = rel_delay_band(flo, flo + width);
		return (int)std::ceil(frac * (max_delay - 1)) + 1;
	}

	// Returns false if the plan is internally inconsistent (a source row
	// index outside its parent band).
	bool build(int nchan_, int max_delay_, double f0_, double df_, double exponent_) {
		reverse_band = df_ < 0;
		if( reverse_band ) { f0_ += (nchan_-1)*df_; df_ = -df_; }
		nchan = nchan_; max_delay = max_delay_; f0 = f0_; df = df_; exponent = exponent_;
		bands.clear(); rows.clear();

		// ---- step 0: one band per channel, half a channel below centre
		bands.emplace_back();
		for( int c=0; c<nchan; ++c ) {
			FdmtBand b;
			b.chan0 = c; b.nchan = 1; b.parent0 = b.parent1 = -1;
			b.ndelay = band_ndelay(chan_freq(c) - 0.5*df, df);
			b.row0 = c ? bands[0][c-1].row0 + bands[0][c-1].ndelay : 0;
			bands[0].push_back(b);
		}
		// ---- merge steps: pair neighbours; odd counts alternate which end
		//      keeps a singleton (ref: src/fdmt.cu:366-387)
		while( bands.back().size() > 1 ) {
			int step = (int)bands.size();
			std::vector<FdmtBand> const& prev = bands[step-1];
			int nprev = (int)prev.size();
			bool odd = nprev % 2;
			bool orphan_first = odd && ((step-1) % 2);
			std::vector<FdmtBand> cur;
			for( int i = orphan_first ? -1 : 0; i < nprev; i += 2 ) {
				FdmtBand b;
				b.parent0 = i;
				b.parent1 = (i+1 < nprev) ? i+1 : -1;
				int n0 = b.parent0 >= 0 ? prev[b.parent0].nchan : 0;
				int n1 = b.parent1 >= 0 ? prev[b.parent1].nchan : 0;
				b.nchan = n0 + n1;
				b.chan0 = cur.empty() ? 0 : cur.back().chan0 + cur.back().nchan;
				b.ndelay = band_ndelay(chan_freq(b.chan0), df * (b.nchan - 1));
				b.row0 = cur.empty() ? 0 : cur.back().row0 + cur.back().ndelay;
				cur.push_back(b);
			}
			bands.push_back(cur);
		}
		nrow_max = 0;
		for( int s=0; s<nstep(); ++s ) nrow_max = std::max(nrow_max, nrow(s));

		// ---- row tables (ref: src/fdmt.cu:446-526)
		bool ok = true;
		rows.resize(nstep());
		for( int s=1; s<nstep(); ++s ) {
			std::vector<FdmtBand> const& prev = bands[s-1];
			rows[s].resize(nrow(s));
			for( FdmtBand const& b : bands[s] ) {
				FdmtBand const* p0 = b.parent0 >= 0 ? &prev[b.parent0] : nullptr;
				FdmtBand const* p1 = b.parent1 >= 0 ? &prev[b.parent1] : nullptr;
				// An absent parent counts as one channel wide, sitting where
				// the present one starts (p0 absent) or ends (p1 absent).
				int n0 = p0 ? p0->nchan : 1;
				int n1 = p1 ? p1->nchan : 1;
				int c0 = (p0 ? p0 : p1)->chan0;
				int c1 = (p1 ? p1 : p0)->chan0;
				if( !p1 ) c1 += n0 - 1;
				double flo    = chan_freq(c0);
				double fmidlo = chan_freq(c0 + n0 - 1);
				double fmidhi = chan_freq(c1);
				double fhi    = chan_freq(c1 + n1 - 1);
				double cmidlo = rel_delay(flo, fmidlo, flo, fhi);
				double cmidhi = rel_delay(flo, fmidhi, flo, fhi);
				for( int d=0; d<b.ndelay; ++d ) {
					int dlo  = (int)std::round(d * cmidlo);
					int dhi  = (int)std::round(d * cmidhi);
					int rest = d - dhi;
					// Keep the high-band row inside its parent
					// (ref: src/fdmt.cu:500-503).
					if( p1 && rest >= p1->ndelay ) rest -= 1;
					if( (p0 && (dlo < 0 || dlo >= p0->ndelay)) ||
					    (p1 && (rest < 0 || rest >= p1->ndelay)) ) ok = false;
					FdmtRow& r = rows[s][b.row0 + d];
					r.src0  = p0 ? p0->row0 + dlo  : -1;
					r.src1  = p1 ? p1->row0 + rest : -1;
					r.delay = dhi;
				}
			}
		}
		return ok;
	}
};

} // namespace bfb
