"""bf.linalg.LinAlg (mirrors python/bifrost/linalg.py:38-67 -> bfLinAlg*)."""
from bifrost_b200.libbifrost import _bf, _check, BifrostObject
from bifrost_b200.ndarray import asarray


class LinAlg(BifrostObject):
    def __init__(self):
        BifrostObject.__init__(self, _bf.bfLinAlgCreate, _bf.bfLinAlgDestroy)

    def matmul(self, alpha, a, b, beta, c):
        """c = alpha*a.b + beta*c; b None: alpha*a.a^H + beta*c; a None:
        alpha*b^H.b + beta*c.  Batch dims follow numpy.matmul."""
        alpha = 1. if alpha is None else float(alpha)
        beta = 0. if beta is None else float(beta)
        a_array = asarray(a).as_BFarray() if a is not None else None
        b_array = asarray(b).as_BFarray() if b is not None else None
        _check(_bf.bfLinAlgMatMul(self.obj, alpha, a_array, b_array, beta,
                                  asarray(c).as_BFarray()))
        return c
