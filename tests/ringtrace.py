"""Scripted use of a bfRing* implementation through ctypes, for differential
tests: the same script of calls runs against libbifrost_b200.so and against
the reference's own ring (oracle/_ref/libbifrost_ref_ring.so, built from the
unmodified src/ring*.cpp by oracle/ref_ring_build.sh), and everything a caller
can observe -- status codes, span sizes and offsets, geometry, sequence
headers, the bytes read back -- is recorded and compared.

`generate()` builds a random script *adaptively* against a live library and
only issues calls that cannot block (it models head / sequence state from the
answers), so a script recorded on the reference replays on any conforming
implementation without threads.
"""
import ctypes
import os
import random
import zlib

from bifrost_b200.libbifrost import _PROTOTYPES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_RING_PATH = os.path.join(ROOT, 'oracle', '_ref', 'libbifrost_ref_ring.so')
NO_TIME_TAG = 2 ** 64 - 1


class _Lib(object):
    pass


def bind(path_or_cdll):
    """ctypes namespace with the ring / proclog / affinity prototypes set."""
    lib = path_or_cdll
    if isinstance(lib, str):
        lib = ctypes.CDLL(lib, mode=os.RTLD_LOCAL | getattr(os, 'RTLD_DEEPBIND', 0))
    ns = _Lib()
    ns._cdll = lib
    for name, (res, args) in _PROTOTYPES.items():
        if not (name.startswith('bfRing') or name.startswith('bfProcLog') or name.startswith('bfAffinity')
                or name in ('bfGetStatusString', 'bfGetAlignment')):
            continue
        fn = getattr(lib, name, None)
        if fn is None:
            continue
        fn.restype, fn.argtypes = res, args
        setattr(ns, name, fn)
    return ns


def load_reference():
    if not os.path.exists(REF_RING_PATH):
        return None
    try:
        return bind(REF_RING_PATH)
    except OSError:
        return None


def load_ours():
    from bifrost_b200.libbifrost import _lib
    return bind(_lib)


def _pattern(seed, nbyte):
    """Deterministic payload: a function of (seed, byte index) only."""
    rnd = random.Random(seed)
    return bytes(rnd.getrandbits(8) for _ in range(nbyte))


class Session(object):
    """One ring plus the handles a script refers to by key."""

    def __init__(self, lib, name=b'trace_ring', space=1):
        self.lib = lib
        self.ring = ctypes.c_void_p()
        st = lib.bfRingCreate(ctypes.byref(self.ring), name, space)
        assert st == 0, st
        self.wseq = None
        self.rseqs, self.wspans, self.rspans = {}, {}, {}

    def close(self):
        for k in list(self.rspans):
            self.lib.bfRingSpanRelease(self.rspans.pop(k))
        for k in list(self.rseqs):
            self.lib.bfRingSequenceClose(self.rseqs.pop(k))
        self.lib.bfRingDestroy(self.ring)

    # ---- helpers
    def _span_info(self, span):
        lib = self.lib
        out = []
        for fn in (lib.bfRingSpanGetSize, lib.bfRingSpanGetOffset, lib.bfRingSpanGetStride,
                   lib.bfRingSpanGetNRinglet):
            v = ctypes.c_ulong()
            assert fn(span, ctypes.byref(v)) == 0
            out.append(int(v.value))
        return out

    def _span_ptr(self, span):
        p = ctypes.c_void_p()
        assert self.lib.bfRingSpanGetData(span, ctypes.byref(p)) == 0
        return p.value

    def _seq_info(self, seq):
        lib = self.lib
        name = ctypes.c_char_p()
        assert lib.bfRingSequenceGetName(seq, ctypes.byref(name)) == 0
        tt = ctypes.c_ulonglong()
        assert lib.bfRingSequenceGetTimeTag(seq, ctypes.byref(tt)) == 0
        hs = ctypes.c_ulong()
        assert lib.bfRingSequenceGetHeaderSize(seq, ctypes.byref(hs)) == 0
        hp = ctypes.c_void_p()
        assert lib.bfRingSequenceGetHeader(seq, ctypes.byref(hp)) == 0
        hdr = ctypes.string_at(hp.value, hs.value) if hs.value else b''
        nr = ctypes.c_ulong()
        assert lib.bfRingSequenceGetNRinglet(seq, ctypes.byref(nr)) == 0
        return [name.value.decode(), int(tt.value), hdr.hex(), int(nr.value)]

    # ---- one call of the script; returns what the caller could observe
    def run(self, op):
        lib, kind = self.lib, op[0]
        if kind == 'resize':
            return [lib.bfRingResize(self.ring, op[1], op[2], op[3])]
        if kind == 'geometry':
            assert lib.bfRingLock(self.ring) == 0
            out = []
            for fn in (lib.bfRingLockedGetContiguousSpan, lib.bfRingLockedGetTotalSpan,
                       lib.bfRingLockedGetNRinglet, lib.bfRingLockedGetStride):
                v = ctypes.c_ulong()
                assert fn(self.ring, ctypes.byref(v)) == 0
                out.append(int(v.value))
            assert lib.bfRingUnlock(self.ring) == 0
            return out
        if kind == 'begin_writing':
            return [lib.bfRingBeginWriting(self.ring)]
        if kind == 'end_writing':
            st = lib.bfRingEndWriting(self.ring)
            ended = ctypes.c_int()
            assert lib.bfRingWritingEnded(self.ring, ctypes.byref(ended)) == 0
            return [st, int(ended.value)]
        if kind == 'seq_begin':
            _, name, tt, hdr, nringlet, off = op
            hdr = bytes.fromhex(hdr)
            h = ctypes.c_void_p()
            buf = ctypes.create_string_buffer(hdr, len(hdr)) if hdr else None
            st = lib.bfRingSequenceBegin(ctypes.byref(h), self.ring, name.encode(), tt, len(hdr),
                                         ctypes.cast(buf, ctypes.c_void_p) if buf else None, nringlet, off)
            if st == 0:
                self.wseq = h
                return [st] + self._seq_info(h)
            return [st]
        if kind == 'seq_end':
            st = lib.bfRingSequenceEnd(self.wseq, op[1])
            self.wseq = None
            return [st]
        if kind == 'reserve':
            _, key, size, nonblocking = op
            h = ctypes.c_void_p()
            st = lib.bfRingSpanReserve(ctypes.byref(h), self.ring, size, nonblocking)
            if st == 0:
                self.wspans[key] = h
                return [st] + self._span_info(h)
            return [st]
        if kind == 'write':
            _, key, seed = op
            h = self.wspans[key]
            size, _, stride, nringlet = self._span_info(h)
            base = self._span_ptr(h)
            for r in range(nringlet):
                ctypes.memmove(base + r * stride, _pattern(seed * 131 + r, size), size)
            return []
        if kind == 'commit':
            _, key, size = op
            return [lib.bfRingSpanCommit(self.wspans.pop(key), size)]
        if kind == 'open':
            _, key, how, arg, guarantee = op
            h = ctypes.c_void_p()
            if how == 'name':
                st = lib.bfRingSequenceOpen(ctypes.byref(h), self.ring, arg.encode(), guarantee)
            elif how == 'at':
                st = lib.bfRingSequenceOpenAt(ctypes.byref(h), self.ring, arg, guarantee)
            elif how == 'latest':
                st = lib.bfRingSequenceOpenLatest(ctypes.byref(h), self.ring, guarantee)
            else:
                st = lib.bfRingSequenceOpenEarliest(ctypes.byref(h), self.ring, guarantee)
            if st == 0:
                self.rseqs[key] = h
                return [st] + self._seq_info(h)
            return [st]
        if kind == 'next':
            st = lib.bfRingSequenceNext(self.rseqs[op[1]])
            return [st] + (self._seq_info(self.rseqs[op[1]]) if st == 0 else [])
        if kind == 'close':
            return [lib.bfRingSequenceClose(self.rseqs.pop(op[1]))]
        if kind == 'acquire':
            _, skey, key, offset, size = op
            h = ctypes.c_void_p()
            st = lib.bfRingSpanAcquire(ctypes.byref(h), self.rseqs[skey], offset, size)
            if st != 0:
                return [st]
            self.rspans[key] = h
            info = self._span_info(h)
            base = self._span_ptr(h)
            crcs = [zlib.crc32(ctypes.string_at(base + r * info[2], info[0])) if info[0] else 0
                    for r in range(info[3])]
            return [st] + info + crcs
        if kind == 'overwritten':
            v = ctypes.c_ulong()
            st = lib.bfRingSpanGetSizeOverwritten(self.rspans[op[1]], ctypes.byref(v))
            return [st, int(v.value)]
        if kind == 'release':
            return [lib.bfRingSpanRelease(self.rspans.pop(op[1]))]
        raise ValueError(kind)


def replay(lib, script, name=b'trace_ring'):
    s = Session(lib, name)
    try:
        return [s.run(tuple(op)) for op in script]
    finally:
        s.close()


def generate(lib, seed, nstep=300, name=b'trace_ring'):
    """Random script + the trace `lib` gave for it.  Never issues a call that
    would block on a conforming implementation."""
    rnd = random.Random(seed)
    s = Session(lib, name)
    script, trace = [], []

    def do(*op):
        out = s.run(op)
        script.append(list(op))
        trace.append(out)
        return out

    gulp = rnd.choice([64, 100, 4096, 5000])
    nringlet = rnd.choice([1, 1, 2, 3])
    do('resize', gulp, gulp * rnd.choice([2, 3, 4]), nringlet)
    do('geometry')
    do('begin_writing')
    head = 0                          # bytes committed so far
    seqs = []                         # [name, begin, finished?, time_tag]
    open_w = []                       # (key, begin, size) in reservation order
    readers = {}                      # key -> index into seqs
    ended = False
    nkey = 0
    for _ in range(nstep):
        can = []
        writing = bool(seqs) and not seqs[-1][2]
        if not ended:
            if not writing and not open_w:
                can += ['seq_begin'] * 3
            if writing:
                can += ['reserve'] * 6
                if not open_w:
                    can += ['seq_end']
            if open_w:
                can += ['commit'] * 6
            if not open_w and not s.rspans:
                can += ['resize']
            if not writing and not open_w and seqs and rnd.random() < 0.03:
                can += ['end_writing']
            if rnd.random() < 0.05:
                can += ['seq_begin']          # may be refused: INVALID_STATE / INVALID_ARGUMENT
        if seqs:
            can += ['open_name', 'open_at']
            if writing or ended:
                can += ['open_edge']
        for key, idx in readers.items():
            can += ['acquire'] * 2
            if idx + 1 < len(seqs) or ended:
                can += ['next']
            can += ['close'] if rnd.random() < 0.2 else []
        if s.rspans:
            can += ['release'] * 3 + ['overwritten']
        if not can:
            break
        what = rnd.choice(can)
        if what == 'seq_begin':
            nm = 'seq%d' % len(seqs) if rnd.random() < 0.9 else rnd.choice(['', 'seq0'])
            # (an unnamed sequence always gets a time tag, so every accepted
            # sequence can be told from the others by what the library reports)
            tt = (len(seqs) + 1) * 1000 if (rnd.random() < 0.8 or not nm) else NO_TIME_TAG
            hdr = _pattern(seed + len(seqs), rnd.choice([0, 5, 40])).hex()
            out = do('seq_begin', nm, tt, hdr, rnd.randint(1, nringlet), 0)
            if out[0] == 0:
                seqs.append([nm, head, False, tt])
        elif what == 'seq_end':
            do('seq_end', 0)
            seqs[-1][2] = True
        elif what == 'reserve':
            if len(open_w) >= 2:
                continue
            size = rnd.choice([gulp, gulp, gulp // 2, 1, gulp + 1 if rnd.random() < 0.1 else gulp])
            nkey += 1
            out = do('reserve', 'w%d' % nkey, size, 1)
            if out[0] == 0:
                open_w.append(('w%d' % nkey, out[2], size))
                do('write', 'w%d' % nkey, nkey)
        elif what == 'commit':
            if len(open_w) == 2 and rnd.random() < 0.2:
                key, _, size = open_w.pop()                    # give the newest back unused
                do('commit', key, 0)
            else:
                key, _, size = open_w.pop(0)
                # a short commit is only legal when nothing is reserved behind it
                csize = size if open_w else rnd.choice([size, size, size // 2, 0])
                out = do('commit', key, csize)
                if out[0] == 0:
                    head += csize
        elif what == 'resize':
            gulp = gulp if rnd.random() < 0.5 else gulp * 2
            do('resize', gulp, gulp * rnd.choice([2, 4, 8]), nringlet)
            do('geometry')
        elif what == 'end_writing':
            out = do('end_writing')
            ended = out[0] == 0
        elif what in ('open_name', 'open_at', 'open_edge'):
            nkey += 1
            key = 'r%d' % nkey
            g = rnd.choice([0, 1])
            if what == 'open_name':
                idx = rnd.randrange(len(seqs))
                out = do('open', key, 'name', seqs[idx][0] or 'nameless', g)
            elif what == 'open_at':
                out = do('open', key, 'at', rnd.choice([500, 1000, 1500, len(seqs) * 1000 + 7]), g)
            else:
                out = do('open', key, rnd.choice(['latest', 'earliest']), 0, g)
            if out[0] == 0:
                match = [i for i, q in enumerate(seqs) if q[0] == out[1] and q[3] == out[2]]
                readers[key] = match[-1]
        elif what == 'acquire':
            key = rnd.choice(sorted(readers))
            nm, begin, finished, _ = seqs[readers[key]]
            have = head - begin
            size = rnd.choice([gulp, gulp // 2, 1])
            if finished:
                off = rnd.choice([0, max(have - size, 0), have, have + 3, rnd.randint(0, max(have, 1))])
            elif have >= size:
                off = rnd.randint(0, have - size)
            else:
                continue
            nkey += 1
            out = do('acquire', key, 'a%d' % nkey, off, size)
        elif what == 'next':
            key = rnd.choice([k for k, i in readers.items() if i + 1 < len(seqs) or ended])
            out = do('next', key)
            if out[0] == 0:
                readers[key] += 1
        elif what == 'close':
            key = rnd.choice(sorted(readers))
            for a in list(s.rspans):
                do('release', a)
            do('close', key)
            del readers[key]
        elif what == 'release':
            do('release', rnd.choice(sorted(s.rspans)))
        elif what == 'overwritten':
            do('overwritten', rnd.choice(sorted(s.rspans)))
    # wind down so that nothing is left open
    for a in list(s.rspans):
        do('release', a)
    while open_w:
        key, _, size = open_w.pop(0)
        do('commit', key, size)
    s.close()
    return script, trace
