"""bf.reduce (mirrors python/bifrost/reduce.py:38-47 -> bfReduce)."""
from bifrost_b200.libbifrost import _bf, _th, _check
from bifrost_b200.ndarray import asarray


def reduce(idata, odata, op='sum'):
    """Reduce exactly one axis of `idata` by the integer factor implied by
    `odata.shape`.  op: sum, mean, min, max, stderr, pwrsum, pwrmean, pwrmin,
    pwrmax, pwrstderr."""
    try:
        op = getattr(_th.BFreduce_enum, op)
    except AttributeError:
        raise ValueError("Invalid reduce op: " + str(op))
    _check(_bf.bfReduce(asarray(idata).as_BFarray(), asarray(odata).as_BFarray(), int(op)))
    return odata
