"""Header-only views (mirrors python/bifrost/views/basic_views.py): they change
how a stream's tensor is described without touching the data."""
from bifrost_b200.DataType import DataType
from bifrost_b200.pipeline import block_view
from bifrost_b200.units import convert_units


def custom(block, hdr_transform):
    return block_view(block, hdr_transform)


def _axis(tensor, axis):
    return tensor['labels'].index(axis) if isinstance(axis, str) else axis


def rename_axis(block, old, new):
    def transform(hdr):
        t = hdr['_tensor']
        t['labels'][t['labels'].index(old)] = new
        return hdr
    return block_view(block, transform)


def reinterpret_axis(block, axis, label, scale=None, units=None):
    def transform(hdr):
        t = hdr['_tensor']
        ax = _axis(t, axis)
        if label is not None:
            t['labels'][ax] = label
        if scale is not None:
            t['scales'][ax] = scale
        if units is not None:
            t['units'][ax] = units
        return hdr
    return block_view(block, transform)


def reverse_scale(block, axis):
    """Flip the sign of the step of `axis` (label or index).  The reference
    (views/basic_views.py:67-75) only acts when the axis is given by label; an
    index works here too."""
    def transform(hdr):
        t = hdr['_tensor']
        ax = _axis(t, axis)
        start, step = t['scales'][ax]
        t['scales'][ax] = [start, -step]
        return hdr
    return block_view(block, transform)


def add_axis(block, axis, label=None, scale=None, units=None):
    def transform(hdr):
        t = hdr['_tensor']
        ax = axis
        if isinstance(ax, str):
            ax = t['labels'].index(ax) + 1
        if ax < 0:
            ax += len(t['shape']) + 1
        t['shape'].insert(ax, 1)
        for key, val in (('labels', label), ('scales', scale), ('units', units)):
            if key in t:
                t[key].insert(ax, val)
        return hdr
    return block_view(block, transform)


def delete_axis(block, axis):
    def transform(hdr):
        t = hdr['_tensor']
        ax = _axis(t, axis)
        if ax < 0:
            ax += len(t['shape'])
        if t['shape'][ax] != 1:
            raise ValueError(f"Cannot delete non-unitary axis {axis} with shape {t['shape'][ax]}")
        for key in ('shape', 'labels', 'scales', 'units'):
            if key in t:
                del t[key][ax]
        return hdr
    return block_view(block, transform)


def astype(block, dtype):
    def transform(hdr):
        t = hdr['_tensor']
        old = DataType(t['dtype']).itemsize * t['shape'][-1]
        new = DataType(dtype).itemsize
        if old % new:
            raise ValueError("New type not compatible with data shape")
        t['shape'][-1] = old // new
        t['dtype'] = str(DataType(dtype))
        return hdr
    return block_view(block, transform)


def split_axis(block, axis, n, label=None):
    def transform(hdr):
        t = hdr['_tensor']
        ax = _axis(t, axis)
        shape = t['shape']
        if shape[ax] == -1:
            # n frames become one (basic_views.py:154-158 of the reference); the
            # reader of this view counts in the new frames (ring.Reader)
            hdr['gulp_nframe'] = (hdr.get('gulp_nframe', 1) - 1) // n + 1
        else:
            if shape[ax] % n:
                raise ValueError(f"Split does not evenly divide axis ({shape[ax]} // {n})")
            shape[ax] //= n
        shape.insert(ax + 1, n)
        if 'units' in t:
            t['units'].insert(ax + 1, t['units'][ax])
        if 'labels' in t:
            t['labels'].insert(ax + 1, label if label is not None else t['labels'][ax] + "_split")
        if 'scales' in t:
            t['scales'].insert(ax + 1, [0, t['scales'][ax][1]])
            t['scales'][ax] = [t['scales'][ax][0], t['scales'][ax][1] * n]
        return hdr
    return block_view(block, transform)


def merge_axes(block, axis1, axis2, label=None):
    def transform(hdr):
        t = hdr['_tensor']
        a1, a2 = sorted([_axis(t, axis1), _axis(t, axis2)])
        if a2 != a1 + 1:
            raise ValueError("Merge axes must be adjacent")
        n = t['shape'][a2]
        if n == -1:
            raise ValueError("Second merge axis cannot be frame axis")
        if t['shape'][a1] == -1:
            hdr['gulp_nframe'] = hdr.get('gulp_nframe', 1) * n      # one frame becomes n
        else:
            t['shape'][a1] *= n
        del t['shape'][a2]
        if 'scales' in t and 'units' in t:
            s1, s2 = t['scales'][a1][1], t['scales'][a2][1]
            s2 = convert_units(s2, t['units'][a2], t['units'][a1])
            if abs(s1 - n * s2) > 1e-8 * max(abs(s1), abs(n * s2), 1e-300):
                raise ValueError(f"Scales of merge axes do not line up: {s1} != {n * s2}")
            t['scales'][a1] = [t['scales'][a1][0], s2]
            del t['scales'][a2]
            del t['units'][a2]
        if 'labels' in t:
            if label is not None:
                t['labels'][a1] = label
            del t['labels'][a2]
        return hdr
    return block_view(block, transform)
