"""A small in-process executor with the block API of the reference's
``bifrost.pipeline`` (python/bifrost/pipeline.py): ``SourceBlock`` /
``TransformBlock`` / ``SinkBlock`` with ``on_sequence`` / ``on_data``,
``block_view``, ``block_scope`` and ``Pipeline.run()``.

Scope note: the reference runs one OS thread per block connected
by ring buffers (pipeline.py:249-261, ring2.py).  That scheduler is host code
outside the GPU hot path and is not rebuilt here; this executor runs the same
block graph synchronously, gulp by gulp, on the calling thread, handing each
block ``ispan.data`` / ``ospan.data`` arrays exactly as the reference does:
  * headers are dicts with a ``_tensor`` entry (dtype, shape with -1 on the
    frame axis, labels, scales, units) -- ring2.py:229-255;
  * spans expose ``.data`` (bf.ndarray in the ring's space), ``.nframe``,
    ``.frame_offset``, ``.tensor``;
  * ``define_input_overlap_nframe`` re-presents the tail of each gulp to the
    next one (blocks/fdmt.py:112-115) and ``on_data`` may return the number of
    frames to commit (blocks/accumulate.py:63-74).
One stream synchronisation per gulp per block, as in pipeline.py:628.
"""
import threading
from copy import deepcopy

import numpy as np

from bifrost_b200 import device
from bifrost_b200.DataType import DataType
from bifrost_b200.ndarray import ndarray, empty, copy_array, memset_array
from bifrost_b200.memory import space_accessible

_tls = threading.local()


def get_default_pipeline():
    if not hasattr(_tls, 'pipeline_stack'):
        _tls.pipeline_stack = [Pipeline()]
    return _tls.pipeline_stack[-1]


def _scope_stack():
    if not hasattr(_tls, 'scope_stack'):
        _tls.scope_stack = [{}]
    return _tls.scope_stack


class block_scope(object):
    """Attribute scope inherited by blocks created inside it (gulp_nframe,
    buffer_nframe, buffer_factor, core, gpu, share_temp_storage, fuse) --
    pipeline.py:84-134.  ``gpu`` selects the device of the blocks' thread;
    ``fuse=True`` collapses chains that have a fused kernel into one launch
    (Pipeline._fuse_chains); ``core``, ``buffer_factor`` and
    ``share_temp_storage`` configure the reference's thread-per-block ring
    scheduler, which this synchronous executor does not have: they are accepted
    and have no effect here."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs

    def __enter__(self):
        merged = dict(_scope_stack()[-1])
        merged.update(self.kwargs)
        _scope_stack().append(merged)
        return self

    def __exit__(self, *exc):
        _scope_stack().pop()


class Pipeline(object):
    def __init__(self, name=None, **kwargs):
        self.name = name or 'Pipeline'
        self.blocks = []
        self.kwargs = kwargs

    def __enter__(self):
        if not hasattr(_tls, 'pipeline_stack'):
            _tls.pipeline_stack = [Pipeline()]
        _tls.pipeline_stack.append(self)
        return self

    def __exit__(self, *exc):
        _tls.pipeline_stack.pop()

    def run(self):
        self._fuse_chains()
        sources = [b for b in self.blocks if isinstance(b, SourceBlock)]
        for src in sources:
            src._run()

    def _fuse_chains(self):
        """Honours ``block_scope(fuse=True)`` (pipeline.py:84-134 of the reference,
        where the flag only removes the ring between blocks): a chain
            transpose(['time','pol','freq','fine_time']) -> fft('fine_time', apply_fftshift=True)
            -> detect('stokes') -> [merge_axes('freq', ...)] -> reduce('freq', f) -> accumulate(n)
        whose blocks were all created under fuse=True, each feeding only the next,
        is replaced by one SpectrometerBlock (one kernel launch per gulp,
        bfSpectrometerFused) between the chain's input and output rings."""
        from bifrost_b200.blocks.transpose import TransposeBlock
        from bifrost_b200.blocks.fft import FftBlock
        from bifrost_b200.blocks.detect import DetectBlock
        from bifrost_b200.blocks.reduce import ReduceBlock
        from bifrost_b200.blocks.accumulate import AccumulateBlock
        from bifrost_b200.blocks.spectrometer import SpectrometerBlock

        def only_consumer(ring, through_views=False):
            """The single block reading `ring` (optionally through one header view)."""
            if len(ring.consumers) == 1 and not ring.views:
                return ring.consumers[0]
            if through_views and not ring.consumers and len(ring.views) == 1:
                return only_consumer(ring.views[0])
            return None

        for t in [b for b in self.blocks if isinstance(b, TransposeBlock) and b.fuse]:
            if list(t.specified_axes) != ['time', 'pol', 'freq', 'fine_time']:
                continue
            f = only_consumer(t.orings[0])
            if not (isinstance(f, FftBlock) and f.fuse and f.specified_axes == ['fine_time'] and f.apply_fftshift
                    and not f.inverse and not f.real_output):
                continue
            d = only_consumer(f.orings[0])
            if not (isinstance(d, DetectBlock) and d.fuse and d.mode == 'stokes' and d.specified_axis in (None, 'pol')):
                continue
            r = only_consumer(d.orings[0], through_views=True)
            if not (isinstance(r, ReduceBlock) and r.fuse and r.specified_axis == 'freq' and r.op == 'sum'
                    and r.specified_factor in (1, 2, 4, 8, 16, 32)):
                continue
            a = only_consumer(r.orings[0])
            if not (isinstance(a, AccumulateBlock) and a.fuse and a.dtype in (None, 'f32')):
                continue
            iring, oring = t.irings[0], a.orings[0]
            if not hasattr(_tls, 'pipeline_stack'):
                _tls.pipeline_stack = [Pipeline()]
            _tls.pipeline_stack.append(self)
            try:
                spec = SpectrometerBlock(iring, f_avg=r.specified_factor, n_int=a.nframe,
                                         gulp_nframe=t.gulp_nframe, gpu=t.gpu, core=t.core)
            finally:
                _tls.pipeline_stack.pop()
            iring.consumers.remove(t)
            spec.orings = [oring]
            oring.owner = spec
            for b in (t, f, d, r, a):
                self.blocks.remove(b)

    def shutdown(self):
        pass


class Ring(object):
    """Stand-in for a ring: remembers its space, owner, consumers and any
    header-only views (``block_view``) hanging off it, and fans sequences and
    spans out to them."""

    def __init__(self, space, owner, header_transform=None):
        self.space = space
        self.owner = owner
        self.consumers = []
        self.views = []
        self.header_transform = header_transform
        self.name = f"ring_{id(self):x}"
        self._seq = None

    def begin(self, hdr):
        self._seq = Sequence(deepcopy(hdr))
        for c in self.consumers:
            c._begin_sequence(self._seq)
        for v in self.views:
            vh = v.header_transform(deepcopy(hdr))
            if vh is None:
                raise ValueError("Header transform returned None")
            v.begin(vh)
        return self._seq

    def push(self, data, frame_offset):
        for c in self.consumers:
            c._push(self._seq, data, frame_offset)
        for v in self.views:
            v.push(v._reinterpret(data), frame_offset)

    def end(self):
        for c in self.consumers:
            c._end_sequence(self._seq)
        for v in self.views:
            v.end()

    def _reinterpret(self, data):
        """Present `data` with this view's tensor shape/dtype (no copy)."""
        tensor = self._seq.tensor
        shape = list(tensor['shape'])
        fax = shape.index(-1)
        known = int(np.prod([s for s in shape if s != -1])) if len(shape) > 1 else 1
        dt = DataType(tensor['dtype'])
        nbyte = data.nbytes
        shape[fax] = (nbyte * 8) // (known * dt.itemsize_bits)
        if list(data.shape) == shape and data.bf.dtype == dt:
            return data
        if not data.flags['C_CONTIGUOUS']:
            raise ValueError("Header views need C-contiguous spans")
        return ndarray(space=data.bf.space, buffer=data.ctypes.data, shape=shape, dtype=dt,
                       native=data.bf.native, conjugated=data.bf.conjugated)


class Sequence(object):
    def __init__(self, header):
        self.header = header
        self.tensor = header['_tensor']
        self.name = header.get('name', '')
        self.time_tag = header.get('time_tag', 0)


class Span(object):
    def __init__(self, sequence, data, frame_offset, frame_axis):
        self.sequence = sequence
        self.tensor = sequence.tensor
        self.data = data
        self.frame_offset = frame_offset
        self.frame_axis = frame_axis
        self.nframe = data.shape[frame_axis]
        self.nframe_skipped = 0
        self.nframe_overwritten = 0
        self.frame_nbyte = 0 if self.nframe == 0 else data.nbytes // max(self.nframe, 1)


def _sync(*spaces):
    """One stream synchronisation per gulp per block (pipeline.py:628), skipped
    when every ring involved lives in plain system memory."""
    if any(str(sp) != 'system' for sp in spaces):
        device.stream_synchronize()


def _frame_axis(tensor):
    return tensor['shape'].index(-1)


def _alloc(tensor, nframe, space):
    shape = list(tensor['shape'])
    shape[_frame_axis(tensor)] = nframe
    return empty(shape, dtype=tensor['dtype'], space=space)


def _slice_frames(arr, axis, lo, hi):
    sl = [slice(None)] * arr.ndim
    sl[axis] = slice(lo, hi)
    return arr[tuple(sl)]


class Block(object):
    instance_counts = {}

    def __init__(self, irings, name=None, type_=None, **kwargs):
        self.type = type_ or self.__class__.__name__
        cnt = Block.instance_counts.get(self.type, 0)
        Block.instance_counts[self.type] = cnt + 1
        self.name = name or f"{self.type}_{cnt}"
        scope = dict(_scope_stack()[-1])
        scope.update(kwargs)
        self.gulp_nframe = scope.get('gulp_nframe')
        self.buffer_nframe = scope.get('buffer_nframe')
        self.buffer_factor = scope.get('buffer_factor')
        self.core = scope.get('core')
        self.gpu = scope.get('gpu')
        self.fuse = scope.get('fuse', False)
        self.pipeline = get_default_pipeline()
        self.pipeline.blocks.append(self)
        self.irings = [r.orings[0] if hasattr(r, 'orings') else r for r in irings]
        for r in self.irings:
            r.consumers.append(self)
        self.orings = []

    def create_ring(self, space):
        return Ring(space, self)

    def get_temp_storage(self, space):
        return None


class SourceBlock(Block):
    def __init__(self, sourcenames, gulp_nframe, space=None, *args, **kwargs):
        super(SourceBlock, self).__init__([], *args, gulp_nframe=gulp_nframe, **kwargs)
        self.sourcenames = sourcenames
        self.orings = [self.create_ring(space=space or 'system')]
        self._seq_count = 0

    def create_reader(self, sourcename):
        raise NotImplementedError

    def on_sequence(self, reader, sourcename):
        raise NotImplementedError

    def on_data(self, reader, ospans):
        raise NotImplementedError

    def _run(self):
        if self.gpu is not None:
            device.set_device(self.gpu)
        for sourcename in self.sourcenames:
            with self.create_reader(sourcename) as reader:
                ohdrs = self.on_sequence(reader, sourcename)
                for ohdr in ohdrs:
                    ohdr.setdefault('time_tag', self._seq_count)
                    ohdr.setdefault('name', f"unnamed-sequence-{self._seq_count}")
                    ohdr.setdefault('gulp_nframe', self.gulp_nframe)
                self._seq_count += 1
                ring = self.orings[0]
                seq = ring.begin(ohdrs[0])
                fax = _frame_axis(seq.tensor)
                offset = 0
                while True:
                    data = _alloc(seq.tensor, self.gulp_nframe, ring.space)
                    ospan = Span(seq, data, offset, fax)
                    nframes = self.on_data(reader, [ospan])
                    _sync(ring.space)
                    n = nframes[0]
                    if n > 0:
                        out = data if n == ospan.nframe else _slice_frames(data, fax, 0, n)
                        ring.push(out, offset)
                        offset += n
                    if n == 0 or n < ospan.nframe:
                        break
                ring.end()


class _ConsumerMixin(object):
    """Input-side buffering shared by transform and sink blocks."""

    def _begin_sequence(self, iseq):
        valid = self.define_valid_input_spaces()
        if valid != 'any' and not any(space_accessible(self.irings[0].space, [s]) for s in valid):
            raise ValueError(f"{self.name}: input space '{self.irings[0].space}' not in {valid}")
        if self.gpu is not None:
            device.set_device(self.gpu)
        self._iseq = iseq
        self._ifax = _frame_axis(iseq.tensor)
        self._pending = None            # frames carried between pushes
        self._pending_offset = 0
        self._overlap = 0
        self._open_outputs(iseq)

    def _gulp(self):
        return self.gulp_nframe or self._iseq.header.get('gulp_nframe') or 1

    def _push(self, iseq, data, frame_offset):
        """Cuts the pushed span into gulps of gulp + overlap frames.  Gulps are
        *views* of one buffer walked with a read cursor; only the tail that is
        left over (fewer than gulp + overlap frames) is copied, once per push,
        in front of the next span -- so the work is linear in the frames pushed
        even when the consumer's gulp is much smaller than the producer's."""
        fax = self._ifax
        gulp, ovl = self._gulp(), self._overlap
        if self._pending is None and ovl == 0 and data.shape[fax] == gulp:
            self._process(data, frame_offset)                 # zero-copy fast path
            return
        if self._pending is None:
            buf, self._pending_offset = data, frame_offset
        else:
            n0, n1 = self._pending.shape[fax], data.shape[fax]
            shape = list(self._pending.shape)
            shape[fax] = n0 + n1
            buf = empty(shape, dtype=self._pending.bf.dtype, space=self._pending.bf.space)
            copy_array(_slice_frames(buf, fax, 0, n0), self._pending)
            copy_array(_slice_frames(buf, fax, n0, n0 + n1), data)
        nbuf, pos = buf.shape[fax], 0
        while nbuf - pos >= gulp + ovl:
            self._process(_slice_frames(buf, fax, pos, pos + gulp + ovl), self._pending_offset)
            pos += gulp
            self._pending_offset += gulp
        if pos == nbuf:
            self._pending = None
        elif pos == 0 and buf is not data:
            self._pending = buf
        else:
            # own copy of the tail: `data` belongs to the producer, and a view
            # would keep the whole span alive
            tail = _slice_frames(buf, fax, pos, nbuf)
            keep = empty(tail.shape, dtype=tail.bf.dtype, space=tail.bf.space)
            copy_array(keep, tail)
            self._pending = keep

    def _end_sequence(self, iseq):
        fax = self._ifax
        if self._pending is not None and self._pending.shape[fax] > self._overlap:
            self._process(self._pending, self._pending_offset)      # ragged final gulp
        self._pending = None
        self.on_sequence_end(iseq)
        self._close_outputs()


class TransformBlock(_ConsumerMixin, Block):
    def __init__(self, iring, *args, **kwargs):
        super(TransformBlock, self).__init__([iring], *args, **kwargs)
        self.iring = self.irings[0]
        self.orings = [self.create_ring(space=self.iring.space)]

    # ---- user hooks (same names/meaning as pipeline.py:703-748)
    def define_valid_input_spaces(self):
        return 'any'

    def define_input_overlap_nframe(self, iseq):
        return 0

    def define_output_nframes(self, input_nframe):
        return input_nframe

    def on_sequence(self, iseq):
        raise NotImplementedError

    def on_sequence_end(self, iseq):
        pass

    def on_data(self, ispan, ospan):
        raise NotImplementedError

    def on_skip(self, islice, ospan):
        memset_array(ospan.data, 0)

    # ---- executor internals
    def _open_outputs(self, iseq):
        ohdr = self.on_sequence(iseq)
        ohdr.setdefault('gulp_nframe', iseq.header.get('gulp_nframe'))
        self._overlap = self.define_input_overlap_nframe(iseq)
        ring = self.orings[0]
        self._oseq = Sequence(deepcopy(ohdr))
        self._ofax = _frame_axis(self._oseq.tensor)
        self._ooffset = 0
        self._ohold = None
        ring.begin(ohdr)

    def _process(self, idata, frame_offset):
        ring = self.orings[0]
        ispan = Span(self._iseq, idata, frame_offset, self._ifax)
        onframe = self.define_output_nframes(ispan.nframe)
        # Blocks that commit rarely (accumulate) keep writing the same output
        odata = self._ohold if self._ohold is not None else _alloc(self._oseq.tensor, onframe, ring.space)
        ospan = Span(self._oseq, odata, self._ooffset, self._ofax)
        ncommit = self.on_data(ispan, ospan)
        _sync(ring.space, self.irings[0].space)
        if ncommit is None:
            ooverlap = self.define_output_nframes(self._overlap) if self._overlap else 0
            ncommit = max(onframe - ooverlap, 0)
        if ncommit == 0:
            self._ohold = odata
            return
        self._ohold = None
        out = odata if ncommit == onframe else _slice_frames(odata, self._ofax, 0, ncommit)
        ring.push(out, self._ooffset)
        self._ooffset += ncommit

    def _close_outputs(self):
        self.orings[0].end()


class SinkBlock(_ConsumerMixin, Block):
    def __init__(self, iring, *args, **kwargs):
        super(SinkBlock, self).__init__([iring], *args, **kwargs)
        self.iring = self.irings[0]

    def define_valid_input_spaces(self):
        return 'any'

    def define_input_overlap_nframe(self, iseq):
        return 0

    def on_sequence(self, iseq):
        raise NotImplementedError

    def on_sequence_end(self, iseq):
        pass

    def on_data(self, ispan):
        raise NotImplementedError

    def _open_outputs(self, iseq):
        self.on_sequence(iseq)
        self._overlap = self.define_input_overlap_nframe(iseq)

    def _process(self, idata, frame_offset):
        self.on_data(Span(self._iseq, idata, frame_offset, self._ifax))
        _sync(self.irings[0].space)

    def _close_outputs(self):
        pass


class BlockView(object):
    """What ``block_view`` returns: usable wherever a block is accepted as an
    input (it only carries an output ring)."""

    def __init__(self, ring, base):
        self.orings = [ring]
        self.base = base


def block_view(block, header_transform):
    """A view of `block` whose output header is passed through
    `header_transform(hdr) -> hdr` (pipeline.py:block_view); no data moves."""
    parent = block.orings[0]
    ring = Ring(parent.space, parent.owner, header_transform)
    parent.views.append(ring)
    return BlockView(ring, block)
