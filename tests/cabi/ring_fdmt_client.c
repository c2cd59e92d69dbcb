/* A plain C client of libbifrost_b200.so: what a compiled pipeline written
 * against the reference's headers does for one FDMT block.
 *
 *   ring_fdmt_client <space> <op> <nchan> <ntime> <gulp> <overlap> <max_delay> <f0> <df> <in.bin> <out.bin>
 *
 * The int8 filterbank in.bin ([nchan][ntime], time fastest) is written gulp by
 * gulp into an input ring whose RINGLETS are the channels (time is the ring's
 * byte axis), so every span is a [nchan][gulp + overlap] view whose rows are
 * a ring stride apart -- the non-contiguous layout arrays from a ring have
 * (trap 1).  Each input span goes through
 *   op = fdmt : bfFdmtExecute into a span of an output ring with one ringlet per delay
 *   op = copy : bfMemcpy2D (plumbing check; runs in system space without a GPU)
 * and the committed output spans are read back and appended to out.bin
 * ([nrow][gulp] per gulp, row-major), as blocks/fdmt.py:112-124 of the
 * reference commits gulp frames per gulp + max_delay frames of input.
 *
 * tests/test_cabi.py compares out.bin with the oracle.  Uses only
 * <bifrost/...> headers; compiled with gcc -std=c99.
 */
#include <bifrost/ring.h>
#include <bifrost/memory.h>
#include <bifrost/array.h>
#include <bifrost/cuda.h>
#include <bifrost/fdmt.h>
#include <bifrost/proclog.h>
#include <bifrost/affinity.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(call) do { BFstatus s_ = (call); if( s_ != BF_STATUS_SUCCESS ) { \
	fprintf(stderr, "%s:%d: %s -> %s\n", __FILE__, __LINE__, #call, bfGetStatusString(s_)); exit(2); } } while(0)

/* device work is asynchronous on the calling thread's stream (pipeline.py:628 syncs once per gulp) */
#define SYNC() do { if( space != BF_SPACE_SYSTEM ) CHECK(bfStreamSynchronize()); } while(0)

static BFarray view2d(void* data, BFspace space, BFdtype dtype, long nrow, long ncol, long pitch, long itemsize) {
	BFarray a;
	memset(&a, 0, sizeof(a));
	a.data = data; a.space = space; a.dtype = dtype; a.ndim = 2;
	a.shape[0] = nrow;   a.shape[1] = ncol;
	a.strides[0] = pitch; a.strides[1] = itemsize;
	return a;
}

int main(int argc, char** argv) {
	if( argc != 12 ) { fprintf(stderr, "usage: see the header of %s\n", __FILE__); return 1; }
	BFspace space    = !strcmp(argv[1], "cuda") ? BF_SPACE_CUDA : BF_SPACE_SYSTEM;
	int     do_fdmt  = !strcmp(argv[2], "fdmt");
	long    nchan    = atol(argv[3]), ntime = atol(argv[4]), gulp = atol(argv[5]), overlap = atol(argv[6]);
	long    max_delay = atol(argv[7]);
	double  f0 = atof(argv[8]), df = atof(argv[9]);
	long    nrow_out = do_fdmt ? max_delay : nchan;
	long    osize    = do_fdmt ? 4 : 1;                       /* f32 banks / copied bytes */

	signed char* host_in = (signed char*)malloc((size_t)nchan * ntime);
	FILE* f = fopen(argv[10], "rb");
	if( !f || fread(host_in, 1, (size_t)nchan * ntime, f) != (size_t)nchan * ntime ) { fprintf(stderr, "cannot read %s\n", argv[10]); return 1; }
	fclose(f);
	FILE* fo = fopen(argv[11], "wb");
	if( !fo ) { fprintf(stderr, "cannot write %s\n", argv[11]); return 1; }

	BFproclog log;
	CHECK(bfProcLogCreate(&log, "ring_fdmt_client/perf"));
	int core = -2;
	CHECK(bfAffinityGetCore(&core));

	/* rings: channels (delays) are ringlets, time is the byte axis */
	BFring iring, oring;
	CHECK(bfRingCreate(&iring, "client_in", space));
	CHECK(bfRingCreate(&oring, "client_out", space));
	CHECK(bfRingResize(iring, (BFsize)(gulp + overlap), (BFsize)(4 * (gulp + overlap)), (BFsize)nchan));
	CHECK(bfRingResize(oring, (BFsize)((gulp + overlap) * osize), (BFsize)(4 * (gulp + overlap) * osize), (BFsize)nrow_out));
	CHECK(bfRingBeginWriting(iring));
	CHECK(bfRingBeginWriting(oring));
	BFwsequence iwseq, owseq;
	const char hdr[] = "{\"name\": \"client\"}";
	CHECK(bfRingSequenceBegin(&iwseq, iring, "client", 0, sizeof(hdr), hdr, (BFsize)nchan, 0));
	CHECK(bfRingSequenceBegin(&owseq, oring, "client", 0, sizeof(hdr), hdr, (BFsize)nrow_out, 0));
	BFrsequence irseq, orseq;
	CHECK(bfRingSequenceOpen(&irseq, iring, "client", 1));
	CHECK(bfRingSequenceOpenEarliest(&orseq, oring, 1));
	const void* got_hdr = NULL; BFsize got_size = 0;
	CHECK(bfRingSequenceGetHeader((BFsequence)orseq, &got_hdr));
	CHECK(bfRingSequenceGetHeaderSize((BFsequence)orseq, &got_size));
	if( got_size != sizeof(hdr) || memcmp(got_hdr, hdr, sizeof(hdr)) ) { fprintf(stderr, "header mismatch\n"); return 3; }

	BFfdmt plan = NULL;
	void*  workspace = NULL; BFsize workspace_size = 0;
	if( do_fdmt ) {
		CHECK(bfFdmtCreate(&plan));
		CHECK(bfFdmtInit(plan, (BFsize)nchan, (BFsize)max_delay, f0, df, -2.0, space, NULL, NULL));
	}

	float* host_out = (float*)malloc((size_t)nrow_out * (gulp + overlap) * osize);
	long written = 0, out_off = 0, ngulp = 0;
	for( long t0 = 0; t0 + overlap < ntime; t0 += gulp ) {
		/* source side: the new frames of this span (the first span also brings the overlap) */
		long have = written, want = t0 + gulp + overlap;
		if( want > ntime ) want = ntime;
		while( have < want ) {
			long n = want - have < gulp ? want - have : gulp;
			BFwspan w; BFspan_info wi;
			CHECK(bfRingSpanReserve(&w, iring, (BFsize)n, 0));
			CHECK(bfRingSpanGetInfo((BFspan)w, &wi));
			CHECK(bfMemcpy2D(wi.data, wi.stride, space, host_in + have, (BFsize)ntime, BF_SPACE_SYSTEM, (BFsize)n, (BFsize)nchan));
			SYNC();
			CHECK(bfRingSpanCommit(w, (BFsize)n));
			have += n;
		}
		written = have;
		/* transform side */
		long nin = want - t0;                                  /* frames in this span (ragged at the end) */
		long nout = nin - overlap;                             /* frames it completes */
		BFrspan r; BFspan_info ri;
		CHECK(bfRingSpanAcquire(&r, irseq, (BFoffset)t0, (BFsize)nin));
		CHECK(bfRingSpanGetInfo((BFspan)r, &ri));
		if( (long)ri.size != nin || (long)ri.offset != t0 || (long)ri.nringlet != nchan ) { fprintf(stderr, "bad read span\n"); return 3; }
		BFwspan w; BFspan_info wi;
		CHECK(bfRingSpanReserve(&w, oring, (BFsize)(nin * osize), 0));
		CHECK(bfRingSpanGetInfo((BFspan)w, &wi));
		if( do_fdmt ) {
			BFarray in  = view2d(ri.data, space, BF_DTYPE_I8,  nchan,     nin, (long)ri.stride, 1);
			BFarray out = view2d(wi.data, space, BF_DTYPE_F32, max_delay, nin, (long)wi.stride, 4);
			BFsize need = 0;
			CHECK(bfFdmtExecute(plan, &in, &out, 0, NULL, &need));                 /* size query */
			if( need > workspace_size ) {
				if( workspace ) CHECK(bfFree(workspace, space));
				CHECK(bfMalloc(&workspace, need, space));
				workspace_size = need;
			}
			need = workspace_size;
			CHECK(bfFdmtExecute(plan, &in, &out, 0, workspace, &need));
		} else {
			CHECK(bfMemcpy2D(wi.data, wi.stride, space, ri.data, ri.stride, space, (BFsize)nin, (BFsize)nchan));
		}
		SYNC();
		CHECK(bfRingSpanCommit(w, (BFsize)(nout * osize)));   /* the overlap frames are not complete yet */
		CHECK(bfRingSpanRelease(r));
		/* sink side */
		BFrspan o; BFspan_info oi;
		CHECK(bfRingSpanAcquire(&o, orseq, (BFoffset)out_off, (BFsize)(nout * osize)));
		CHECK(bfRingSpanGetInfo((BFspan)o, &oi));
		if( (long)oi.size != nout * osize ) { fprintf(stderr, "bad output span\n"); return 3; }
		CHECK(bfMemcpy2D(host_out, (BFsize)(nout * osize), BF_SPACE_SYSTEM, oi.data, oi.stride, space, (BFsize)(nout * osize), (BFsize)nrow_out));
		SYNC();
		CHECK(bfRingSpanRelease(o));
		fwrite(&nout, sizeof(long), 1, fo);
		fwrite(host_out, 1, (size_t)nrow_out * nout * osize, fo);
		out_off += nout * osize;
		++ngulp;
		char text[128];
		snprintf(text, sizeof(text), "ngulp : %ld\ncore : %d\n", ngulp, core);
		CHECK(bfProcLogUpdate(log, text));
	}
	fclose(fo);
	CHECK(bfRingSequenceEnd(iwseq, 0));
	CHECK(bfRingSequenceEnd(owseq, 0));
	CHECK(bfRingSequenceClose(irseq));
	CHECK(bfRingSequenceClose(orseq));
	CHECK(bfRingEndWriting(iring));
	CHECK(bfRingEndWriting(oring));
	if( plan ) CHECK(bfFdmtDestroy(plan));
	if( workspace ) CHECK(bfFree(workspace, space));
	CHECK(bfRingDestroy(iring));
	CHECK(bfRingDestroy(oring));
	CHECK(bfProcLogDestroy(log));
	printf("OK %ld gulps\n", ngulp);
	return 0;
}
