// shape.hpp -- strided-view canonicalisation.  Several arrays that share one
// logical shape are reduced together to the fewest dims that still describe
// every array's memory layout (unit dims dropped, mutually contiguous
// neighbours fused).  This is what lets the kernels take <=4-D descriptors
// while the C ABI accepts 8-D BFarrays with padded (ring) strides.
#pragma once
#include "core.hpp"

namespace bfb {

struct StridedView {
	int  ndim;
	long shape[BF_MAX_DIMS];
	long strides[BF_MAX_DIMS];   // bytes
};

inline void load_view(BFarray const* a, StridedView* v) {
	v->ndim = a->ndim;
	for( int d=0; d<a->ndim; ++d ) {
		v->shape[d]   = a->shape[d];
		v->strides[d] = a->strides[d];
	}
}

// All views must have identical shape.  Result always has ndim >= 1.
inline void merge_views(StridedView* v, int nview) {
	int nd = v[0].ndim;
	int o = 0;   // number of output dims so far
	for( int d=0; d<nd; ++d ) {
		long len = v[0].shape[d];
		if( len == 1 && nd > 1 ) continue;   // unit dims carry no layout
		bool fuse = (o > 0);
		if( fuse ) {
			for( int k=0; k<nview; ++k ) {
				if( v[k].strides[o-1] != v[k].strides[d] * len ) { fuse = false; break; }
			}
		}
		if( fuse ) {
			for( int k=0; k<nview; ++k ) {
				v[k].shape[o-1]  *= len;
				v[k].strides[o-1] = v[k].strides[d];
			}
		} else {
			for( int k=0; k<nview; ++k ) {
				v[k].shape[o]   = len;
				v[k].strides[o] = v[k].strides[d];
			}
			++o;
		}
	}
	if( o == 0 ) {   // every dim was 1
		for( int k=0; k<nview; ++k ) {
			v[k].shape[0] = 1;
			v[k].strides[0] = v[k].strides[nd-1];
		}
		o = 1;
	}
	for( int k=0; k<nview; ++k ) v[k].ndim = o;
}

// Alignment (power of two, <= cap) common to a base pointer and all strides
// of dims with extent > 1.
inline unsigned long view_alignment(const void* ptr, StridedView const& v,
                                    unsigned long cap) {
	unsigned long a = pow2_alignment((unsigned long)(uintptr_t)ptr, cap);
	for( int d=0; d<v.ndim; ++d ) {
		if( v.shape[d] > 1 ) {
			long s = v.strides[d] < 0 ? -v.strides[d] : v.strides[d];
			a = pow2_alignment((unsigned long)s, a);
		}
	}
	return a;
}

} // namespace bfb
