"""Import-only shim for the reference's cell.py.

``main_flownetS_pyramid_noprevloss_dataloader.py:7`` imports ``ConvLSTMCell`` / ``ConvGRUCell`` but no
script ever instantiates them (zero call sites), so they are not on the inference hot path.  The
names exist so that a drop-in ``from cell import ConvLSTMCell, ConvGRUCell`` keeps working.
"""


class _NotOnHotPath(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} has no call site on the inference path and is not implemented")


class ConvLSTMCell(_NotOnHotPath):
    pass


class ConvGRUCell(_NotOnHotPath):
    pass
