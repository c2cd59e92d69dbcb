"""Small helpers the blocks share for rewriting sequence headers.  A `_tensor`
entry describes the stream: dtype, shape (-1 marks the frame axis), and per-axis
labels / scales / units (any of which may be absent)."""
from copy import deepcopy

PER_AXIS = ('shape', 'labels', 'scales', 'units')


def derive(ihdr):
    """Output header that starts out as a deep copy of the input's."""
    ohdr = deepcopy(ihdr)
    return ohdr, ohdr['_tensor']


def axis_index(tensor, axis):
    """An axis given by label or by position."""
    return tensor['labels'].index(axis) if isinstance(axis, str) else axis


def frame_axis(tensor):
    return tensor['shape'].index(-1)


def remap_axes(itensor, otensor, picks):
    """otensor's per-axis lists become [itensor[...][i] for i in picks] (a pick
    may repeat: each occurrence gets its own copy)."""
    for key in PER_AXIS:
        if key in itensor:
            otensor[key] = [deepcopy(itensor[key][i]) for i in picks]


def scale_step(tensor, axis, factor):
    """Multiplies the sampling step of `axis` (scales = [origin, step])."""
    if 'scales' in tensor and tensor['scales'][axis] is not None:
        origin, step = tensor['scales'][axis]
        tensor['scales'][axis] = [origin, step * factor]
