"""Shared machinery of the drop-in nn.Modules: packed (BN-folded, kernel-layout) weights are built lazily at the
first forward and dropped whenever parameters may have changed (load_state_dict, .to()/.cuda(), .train())."""
import torch

from . import _native as N
from . import _ops as ops


class NativeModule(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self._packed = None

    # -- cache invalidation ---------------------------------------------------------------------------------
    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._packed = None
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        for m in self.modules():            # nested blocks are loaded through _load_from_state_dict, not this method
            if isinstance(m, NativeModule):
                m._packed = None
        return out

    def train(self, mode=True):
        self._packed = None
        return super().train(mode)

    def invalidate_packed_weights(self):
        """Call after modifying parameters in place."""
        self._packed = None

    def packed(self):
        if self.training:
            raise RuntimeError("%s: the B200 path implements inference only -- call .eval() (BatchNorm is folded into "
                               "the convolution weights)" % type(self).__name__)
        if self._packed is None:
            p = next(self.parameters())
            if not p.is_cuda and not N.DRYRUN:
                raise RuntimeError("%s: parameters are on %s; move the module to a CUDA device (no CPU fallback)"
                                   % (type(self).__name__, p.device))
            with torch.no_grad():
                self._packed = self._pack()
        return self._packed

    def _pack(self):
        raise NotImplementedError


def pack_cbr(seq, stride=1):
    """conv_layer Sequential -> PackedConv (Conv[, BN, ReLU])."""
    has_bn = len(seq) > 1
    return ops.PackedConv(seq[0].weight, None, seq[1] if has_bn else None, stride=stride, act=N.ACT_RELU if has_bn else N.ACT_NONE)


def pack_head(seq):
    return ops.PackedConv(seq[0].weight, seq[0].bias, None, stride=1, act=N.ACT_SIGMOID)
