// Multi-threaded stress of the native ring (csrc/ring.cpp) for ThreadSanitizer
// and for data integrity under contention: one writer (three sequences, full
// and short commits), two guaranteed readers with different span sizes, one
// unguaranteed reader that is allowed to be lapped, a thread that keeps growing
// the ring while spans are in flight, and one that polls the locked getters.
// Every byte read is checked against a function of (sequence, offset, ringlet).
// Prints "OK <bytes>" and exits 0.  TEST INFRASTRUCTURE (tests/test_ring_stress.py).
#include <bifrost/ring.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define CHECK(call) do { BFstatus s_ = (call); if( s_ != BF_STATUS_SUCCESS ) { \
	std::fprintf(stderr, "%s:%d: %s -> %d\n", __FILE__, __LINE__, #call, (int)s_); std::exit(2); } } while(0)

static const long GULP = 1000, NGULP = 600, NSEQ = 3, NRINGLET = 3;

static inline unsigned char pattern(long seq, long offset, long ringlet) {
	return (unsigned char)((offset * 31 + seq * 7 + ringlet * 101 + (offset >> 8)) & 0xFF);
}

static std::atomic<long> g_bad(0), g_read(0), g_lapped(0);
static std::atomic<bool> g_done(false);

static void writer(BFring ring) {
	CHECK(bfRingBeginWriting(ring));
	unsigned rnd = 12345;
	for( long s=0; s<NSEQ; ++s ) {
		BFwsequence seq;
		std::string name = "seq" + std::to_string(s);
		long hdr = s;
		CHECK(bfRingSequenceBegin(&seq, ring, name.c_str(), (BFoffset)(100 + s), sizeof(hdr), &hdr, NRINGLET, 0));
		long offset = 0;
		while( offset < GULP * NGULP ) {
			BFwspan span; BFspan_info info;
			CHECK(bfRingSpanReserve(&span, ring, GULP, 0));
			CHECK(bfRingSpanGetInfo((BFspan)span, &info));
			rnd = rnd * 1664525u + 1013904223u;
			long n = ((rnd >> 16) % 4 == 0) ? GULP / 2 : GULP;           // some short commits
			if( n > GULP * NGULP - offset ) n = GULP * NGULP - offset;
			for( long r=0; r<(long)info.nringlet && r<NRINGLET; ++r ) {
				unsigned char* p = (unsigned char*)info.data + r * info.stride;
				for( long i=0; i<n; ++i ) p[i] = pattern(s, offset + i, r);
			}
			CHECK(bfRingSpanCommit(span, (BFsize)n));
			offset += n;
		}
		CHECK(bfRingSequenceEnd(seq, 0));
	}
	CHECK(bfRingEndWriting(ring));
}

static long count_bad(const BFspan_info& info, long s) {
	long bad = 0;
	for( long r=0; r<NRINGLET; ++r ) {
		const unsigned char* p = (const unsigned char*)info.data + r * info.stride;
		for( long i=0; i<(long)info.size; ++i ) bad += p[i] != pattern(s, (long)info.offset + i, r);
	}
	return bad;
}
// An unguaranteed reader may be lapped: the writer then overwrites what it is
// reading -- by design (it asks bfRingSpanGetSizeOverwritten afterwards and
// discards the span), so those reads are not for the sanitizer to judge.
__attribute__((no_sanitize("thread"), noinline))
static long count_bad_unguarded(const BFspan_info& info, long s) {
	long bad = 0;
	for( long r=0; r<NRINGLET; ++r ) {
		const volatile unsigned char* p = (const volatile unsigned char*)info.data + r * info.stride;
		for( long i=0; i<(long)info.size; ++i ) bad += p[i] != pattern(s, (long)info.offset + i, r);
	}
	return bad;
}

static void reader(BFring ring, long span_size, bool guaranteed) {
	BFrsequence seq;
	BFstatus st = guaranteed ? bfRingSequenceOpenEarliest(&seq, ring, 1) : bfRingSequenceOpenLatest(&seq, ring, 0);
	if( st == BF_STATUS_END_OF_DATA ) return;
	CHECK(st);
	for( ;; ) {
		const void* hdr; CHECK(bfRingSequenceGetHeader((BFsequence)seq, &hdr));
		long s = *(const long*)hdr;
		long offset = 0;
		for( ;; ) {
			BFrspan span;
			st = bfRingSpanAcquire(&span, seq, (BFoffset)offset, (BFsize)span_size);
			if( st == BF_STATUS_END_OF_DATA ) break;
			CHECK(st);
			BFspan_info info;
			CHECK(bfRingSpanGetInfo((BFspan)span, &info));
			if( !guaranteed && (offset / span_size) % 16 == 0 )          // dawdle: get lapped now and then
				std::this_thread::sleep_for(std::chrono::microseconds(300));
			long bad = guaranteed ? count_bad(info, s) : count_bad_unguarded(info, s);
			BFsize lost = 0;
			CHECK(bfRingSpanGetSizeOverwritten(span, &lost));
			if( guaranteed ) { if( bad || lost || (long)info.offset != offset ) g_bad += 1 + bad; }
			else if( lost ) { g_lapped += 1; }                            // the writer got there first: contents are void
			else if( bad )  { g_bad += bad; }
			g_read += (long)info.size * NRINGLET;
			offset = (long)info.offset + (long)info.size + (info.size ? 0 : span_size);
			CHECK(bfRingSpanRelease(span));
		}
		st = bfRingSequenceNext(seq);
		if( st == BF_STATUS_END_OF_DATA ) break;
		CHECK(st);
	}
	CHECK(bfRingSequenceClose(seq));
}

static void grower(BFring ring) {
	BFsize total = 8 * GULP;
	for( int i=0; i<6 && !g_done; ++i ) {
		std::this_thread::sleep_for(std::chrono::milliseconds(15));
		total *= 2;
		CHECK(bfRingResize(ring, 1500 + 500 * i, total, NRINGLET));
	}
}

static void peeker(BFring ring) {
	while( !g_done ) {
		BFsize a, b, c, d; void* p;
		CHECK(bfRingLock(ring));
		CHECK(bfRingLockedGetContiguousSpan(ring, &a));
		CHECK(bfRingLockedGetTotalSpan(ring, &b));
		CHECK(bfRingLockedGetStride(ring, &c));
		CHECK(bfRingLockedGetNRinglet(ring, &d));
		CHECK(bfRingLockedGetData(ring, &p));
		CHECK(bfRingUnlock(ring));
		if( c != a + b || d != (BFsize)NRINGLET || !p ) g_bad += 1;
		std::this_thread::sleep_for(std::chrono::milliseconds(1));
	}
}

int main() {
	BFring ring;
	CHECK(bfRingCreate(&ring, "stress", BF_SPACE_SYSTEM));
	CHECK(bfRingResize(ring, 1500, 8 * GULP, NRINGLET));
	std::vector<std::thread> readers;
	readers.emplace_back(reader, ring, 1000L, true);
	readers.emplace_back(reader, ring, 1500L, true);
	readers.emplace_back(reader, ring, 700L, false);
	std::thread g(grower, ring), p(peeker, ring);
	// the guaranteed readers hold their place before data flows (they wait for the first sequence)
	std::this_thread::sleep_for(std::chrono::milliseconds(50));
	std::thread w(writer, ring);
	w.join();
	for( std::thread& t : readers ) t.join();
	g_done = true;
	g.join(); p.join();
	CHECK(bfRingDestroy(ring));
	if( g_bad ) { std::fprintf(stderr, "BAD %ld\n", (long)g_bad); return 1; }
	std::printf("OK %ld bytes read, unguaranteed reader lapped %ld times\n", (long)g_read, (long)g_lapped);
	return 0;
}
