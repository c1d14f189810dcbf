"""Loss entry points under the reference's names (dvmvs/losses.py; imported by fusionnet/run-training.py:13 and
pairnet/run-training.py).  `update_losses` evaluates every prediction scale in ONE fused launch
(dvmvs.training.multi_scale_depth_loss -> depth_loss_forward_kernel; backward depth_loss_backward_kernel) instead of the
reference's ~25 elementwise ops per scale; the LossMeter bookkeeping is host-side as in the reference (it calls .item(), i.e.
synchronises, exactly where the reference does, losses.py:43-46)."""
import torch

from .training import LOSS_TYPES, multi_scale_depth_loss


class LossMeter(object):
    """losses.py:7-23: running sum / count / average of a loss over a dataset pass."""

    def __init__(self):
        self.count = 0.0
        self.sum = 0.0
        self.avg = 0.0
        self.item_average = 0.0

    def update(self, loss, count):
        self.sum += loss
        self.count += count
        self.avg = self.sum / self.count
        self.item_average = loss / count

    def __repr__(self):
        return '{:.4f} ({:.4f})'.format(self.item_average, self.avg)


def calculate_loss(groundtruth, prediction):
    """losses.py:53-82 for one prediction scale: (l1, huber, l1_inv, l1_rel, valid_count) -- the four sums as 0-dim CUDA
    tensors, the count as a python int (the reference's `valid_mask.nonzero().size()[0]` synchronises too).

    NOT differentiable here: the sums are monitoring values (the fused kernel marks them non-differentiable).  The reference's
    calculate_loss returns sums one can back-propagate through; in this package the optimizer loss comes from
    update_losses(..., is_training=True) / dvmvs.training.multi_scale_depth_loss, which is what run-training.py:269-278 uses.  A
    caller that builds its own loss from these sums gets `does not require grad` from autograd, not silent zeros."""
    _, sums = multi_scale_depth_loss([prediction], [1.0], groundtruth, "L1")
    return sums[0, 0], sums[0, 1], sums[0, 2], sums[0, 3], int(sums[0, 4].item())


def update_losses(predictions, weights, groundtruth, is_training, l1_meter, huber_meter, l1_inv_meter, l1_rel_meter, loss_type):
    """losses.py:26-50.  Training: optimizer loss = sum_j weights[j] * loss_j / valid_count_j over all scales (differentiable);
    the meters record the sums of the LAST prediction (the full-resolution one), as in the reference."""
    if is_training:
        if loss_type not in LOSS_TYPES:
            raise ValueError("loss_type must be one of %s" % sorted(LOSS_TYPES))
        optimizer_loss, sums = multi_scale_depth_loss(list(predictions), list(weights), groundtruth, loss_type)
        last = sums[len(predictions) - 1]
    else:
        optimizer_loss = 0
        with torch.no_grad():
            _, sums = multi_scale_depth_loss([predictions[-1]], [1.0], groundtruth, "L1")
        last = sums[0]
    l1, huber, l1_inv, l1_rel, count = [float(v) for v in last.tolist()]       # one device->host copy for all five
    l1_meter.update(l1, count)
    huber_meter.update(huber, count)
    l1_inv_meter.update(l1_inv, count)
    l1_rel_meter.update(l1_rel, count)
    return optimizer_loss
