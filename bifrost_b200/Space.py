"""Memory-space names <-> BFspace (mirrors python/bifrost/Space.py)."""
from bifrost_b200.libbifrost import _bf, _th, _string2space, _space2string

SPACEMAP_TO_STR = {_bf.BF_SPACE_AUTO: 'auto',
                   _bf.BF_SPACE_SYSTEM: 'system',
                   _bf.BF_SPACE_CUDA: 'cuda',
                   _bf.BF_SPACE_CUDA_HOST: 'cuda_host',
                   _bf.BF_SPACE_CUDA_MANAGED: 'cuda_managed'}
SPACEMAP_FROM_STR = {v: k for k, v in SPACEMAP_TO_STR.items()}


class Space(object):
    def __init__(self, s):
        if isinstance(s, Space):
            self._space = s._space
        elif isinstance(s, str):
            if s not in SPACEMAP_FROM_STR:
                raise ValueError(f"Invalid space: '{s}'. Valid spaces: {list(SPACEMAP_FROM_STR)}")
            self._space = s
        elif int(s) in SPACEMAP_TO_STR:
            self._space = SPACEMAP_TO_STR[int(s)]
        else:
            raise ValueError(f"'{s}' is not a space")

    def as_BFspace(self):
        return SPACEMAP_FROM_STR[self._space]

    def __str__(self):
        return self._space

    def __repr__(self):
        return f"Space('{self._space}')"

    def __eq__(self, other):
        return str(self) == str(Space(other)) if not isinstance(other, Space) else self._space == other._space

    def __hash__(self):
        return hash(self._space)
