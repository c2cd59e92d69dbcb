"""bf.fdmt.Fdmt (mirrors python/bifrost/fdmt.py:38-76 -> bfFdmt*)."""
from bifrost_b200.libbifrost import _bf, _check, _get, BifrostObject
from bifrost_b200.ndarray import asarray
from bifrost_b200.Space import Space


class Fdmt(BifrostObject):
    def __init__(self):
        BifrostObject.__init__(self, _bf.bfFdmtCreate, _bf.bfFdmtDestroy)

    def init(self, nchan, max_delay, f0, df, exponent=-2.0, space='cuda'):
        space = Space(space)
        _check(_bf.bfFdmtInit(self.obj, nchan, max_delay, f0, df, exponent,
                              space.as_BFspace(), None, None))

    def execute(self, idata, odata, negative_delays=False):
        _check(_bf.bfFdmtExecute(self.obj, asarray(idata).as_BFarray(),
                                 asarray(odata).as_BFarray(), negative_delays, None, None))
        return odata

    def get_workspace_size(self, idata, odata, negative_delays=False):
        # the workspace depends on the schedule, and negative_delays takes the
        # step-by-step one: query with the flag the execute call will carry
        return _get(_bf.bfFdmtExecute, self.obj, asarray(idata).as_BFarray(),
                    asarray(odata).as_BFarray(), negative_delays, None)

    def execute_workspace(self, idata, odata, workspace_ptr, workspace_size,
                          negative_delays=False):
        size = _bf.BFsize(workspace_size)
        _check(_bf.bfFdmtExecute(self.obj, asarray(idata).as_BFarray(),
                                 asarray(odata).as_BFarray(), negative_delays,
                                 workspace_ptr, size))
        return odata
