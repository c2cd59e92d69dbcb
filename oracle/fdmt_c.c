/* fdmt_c.c -- C restatement of the reference FDMT execution for the timed CPU
 * baseline.  TEST / BENCHMARK INFRASTRUCTURE ONLY (see oracle/__init__.py): the
 * product never links or loads it.
 *
 * Follows the reference kernels one to one, step by step through two ping-pong
 * buffers, exactly as BFfdmt_impl::execute does (src/fdmt.cu:629-718):
 *   init  src/fdmt.cu:52-92    running means per channel, NaN where t < d
 *   step  src/fdmt.cu:95-155   out[r][t] = A[src0][t] + (t >= delay ? A[src1][t-delay] : 0)
 *   final diagonal store out[r][t-r] for t >= r (src/fdmt.cu:120-124,702-708)
 * The plan tables (row offsets, source rows, delays) come from oracle/fdmt.py.
 * Rows are distributed over pthreads (the reference has no CPU path, so
 * this is the strongest straightforward host implementation of its algorithm).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* Minimal pthread parallel-for (this image's gcc ships without libgomp):
 * rows are dealt to threads in interleaved chunks. */
typedef void (*row_fn)(long lo, long hi, void* ctx);
typedef struct { row_fn fn; void* ctx; long n, chunk; int tid, nthread; } par_arg;
static void* par_worker(void* p) {
	par_arg* a = (par_arg*)p;
	for( long lo = (long)a->tid * a->chunk; lo < a->n; lo += (long)a->nthread * a->chunk ) {
		long hi = lo + a->chunk < a->n ? lo + a->chunk : a->n;
		a->fn(lo, hi, a->ctx);
	}
	return NULL;
}
static void parallel_for(long n, long chunk, row_fn fn, void* ctx, int nthread) {
	if( nthread <= 1 || n <= chunk ) { fn(0, n, ctx); return; }
	if( nthread > 256 ) nthread = 256;
	pthread_t th[256];
	par_arg args[256];
	for( int t=0; t<nthread; ++t ) {
		args[t].fn = fn; args[t].ctx = ctx; args[t].n = n; args[t].chunk = chunk;
		args[t].tid = t; args[t].nthread = nthread;
		pthread_create(&th[t], NULL, par_worker, &args[t]);
	}
	for( int t=0; t<nthread; ++t ) pthread_join(th[t], NULL);
}

static inline float load_in(const void* in, int kind, long idx) {
	switch( kind ) {
	case 0:  return (float)((const int8_t*)in)[idx];
	case 1:  return (float)((const uint8_t*)in)[idx];
	case 2:  return (float)((const int16_t*)in)[idx];
	case 3:  return (float)((const uint16_t*)in)[idx];
	case 4:  return (float)((const int32_t*)in)[idx];
	case 5:  return (float)((const uint32_t*)in)[idx];
	default: return ((const float*)in)[idx];
	}
}

typedef struct {
	const void* in; int in_kind; long nchan, ntime; int reverse_band;
	const long* row_offsets; float* buf;
} init_ctx;
static void init_rows(long c_lo, long c_hi, void* p) {
	init_ctx* k = (init_ctx*)p;
	long ntime = k->ntime;
	for( long c=c_lo; c<c_hi; ++c ) {
		long c_in = k->reverse_band ? k->nchan-1-c : c;
		long ndelay = k->row_offsets[c+1] - k->row_offsets[c];
		for( long t=0; t<ntime; ++t ) {
			float tmp = 0.f;
			for( long d=0; d<ndelay; ++d ) {
				float val = NAN;
				if( t >= d ) {
					tmp += load_in(k->in, k->in_kind, c_in*ntime + (t-d));
					val = tmp * (1.f/(float)(d+1));
				}
				k->buf[(k->row_offsets[c]+d)*ntime + t] = val;
			}
		}
	}
}

typedef struct {
	const float* cur; float* nxt; float* out; long ostride, ntime;
	const long* src; const long* dly; int final;
} step_ctx;
static void step_rows(long r_lo, long r_hi, void* p) {
	step_ctx* k = (step_ctx*)p;
	long ntime = k->ntime;
	for( long r=r_lo; r<r_hi; ++r ) {
		long s0 = k->src[2*r], s1 = k->src[2*r+1], d = k->dly[r];
		const float* a = s0 >= 0 ? k->cur + s0*ntime : NULL;
		const float* b = s1 >= 0 ? k->cur + s1*ntime : NULL;
		if( !k->final ) {
			float* o = k->nxt + r*ntime;
			long tmid = d < ntime ? d : ntime;
			for( long t=0; t<tmid; ++t )     o[t] = a ? a[t] : 0.f;
			for( long t=tmid; t<ntime; ++t ) o[t] = (a ? a[t] : 0.f) + (b ? b[t-d] : 0.f);
		} else {
			float* o = k->out + r*k->ostride;
			for( long t=r; t<ntime; ++t ) {
				float v = a ? a[t] : 0.f;
				if( t >= d && b ) v += b[t-d];
				o[t-r] = v;
			}
		}
	}
}

/* Returns 0 on success.  buf_a / buf_b: nrow_max * ntime floats each. */
int fdmt_c_execute(const void* in, int in_kind, long nchan, long ntime, int reverse_band,
                   const long* row_offsets,          /* [nchan+1] */
                   int nstep, const long* step_nrow, /* [nstep] */
                   const long* const* srcrows,       /* [nstep] -> [nrow][2] */
                   const long* const* delays,        /* [nstep] -> [nrow] */
                   float* out, long ostride, float* buf_a, float* buf_b, int nthread) {
	if( nthread <= 0 ) nthread = (int)sysconf(_SC_NPROCESSORS_ONLN);
	init_ctx ic = {in, in_kind, nchan, ntime, reverse_band, row_offsets, buf_a};
	parallel_for(nchan, 8, init_rows, &ic, nthread);
	float* cur = buf_a;
	float* nxt = buf_b;
	for( int s=1; s<nstep; ++s ) {
		step_ctx sc = {cur, nxt, out, ostride, ntime, srcrows[s], delays[s], s == nstep-1};
		parallel_for(step_nrow[s], 2, step_rows, &sc, nthread);
		float* tmp = cur; cur = nxt; nxt = tmp;
	}
	return 0;
}

int fdmt_c_max_threads(void) {
	return (int)sysconf(_SC_NPROCESSORS_ONLN);
}
