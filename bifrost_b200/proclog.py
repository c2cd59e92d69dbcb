"""Per-process status logs (ref: python/bifrost/proclog.py; C side
csrc/ring.cpp, src/proclog.cpp in the reference).

A ``ProcLog("block/quantity")`` is one small text file of ``key : value``
lines under ``PROCLOG_DIR/<pid>/block/quantity`` that the owner rewrites
whenever the values change; monitoring tools (the reference's ``like_top.py``,
``like_bmon.py`` ...) read them from outside with ``load_by_pid``.  The
executor writes the same logs the reference's does: ``<block>/bind``, ``/in``,
``/out``, ``/sequence0`` and ``/perf`` (acquire_time, reserve_time,
process_time per gulp), and every ring describes itself under ``rings/<name>``.
"""
import os
import time

from bifrost_b200.libbifrost import _bf, _check, BifrostObject

PROCLOG_DIR = os.environ.get('BIFROST_B200_PROCLOG_DIR') or _bf.BF_PROCLOG_DIR


class ProcLog(BifrostObject):
    def __init__(self, name):
        BifrostObject.__init__(self, _bf.bfProcLogCreate, _bf.bfProcLogDestroy, name.encode())

    def update(self, contents):
        """Replaces the log's contents by a string or by the items of a dict."""
        if contents is None:
            raise ValueError("Contents cannot be None")
        if isinstance(contents, dict):
            contents = '\n'.join(f'{key} : {value}' for key, value in contents.items())
        _check(_bf.bfProcLogUpdate(self.obj, contents.encode()))


def _number_or_text(text):
    for convert in (lambda t: int(t, 10), float):
        try:
            return convert(text)
        except ValueError:
            pass
    return text


def load_by_filename(filename):
    """One log file as a dict (numbers converted)."""
    for _ in range(5):                     # the writer truncates before it writes
        if os.path.getsize(filename):
            break
        time.sleep(0.001)
    contents = {}
    with open(filename) as f:
        for line in f.read().split('\n'):
            key, colon, value = line.partition(':')
            if colon:
                contents[key.strip()] = _number_or_text(value.strip())
    return contents


def load_by_pid(pid, include_rings=False):
    """All logs of a process: {block: {log: {key: value}}}."""
    base = os.path.join(PROCLOG_DIR, str(pid))
    if not os.path.isdir(base):
        raise RuntimeError(f"Cannot find log directory associated with PID {pid}")
    contents = {}
    for parent, _, filenames in os.walk(base):
        block = os.path.basename(parent)
        if block == 'rings' and not include_rings:
            continue
        for name in filenames:
            try:
                contents.setdefault(block, {})[name] = load_by_filename(os.path.join(parent, name))
            except (IOError, OSError):
                continue
    return contents
