"""MVSLayernormConvLSTMCell with the reference's constructor / forward signature and parameter name
(`conv.weight`, reference dvmvs/convlstm.py:7-64).  Forward = hidden-state warp kernel (pose algebra, bilinear
gather and <=0.01 mask fused) -> 3x3 gate convolution -> gate epilogue kernel (sigmoid / LayerNorm over (h,w) / CELU /
state update).

The gate convolution over cat[input, hidden] (convlstm.py:43-44) is linear in its input channels, so it is evaluated as
conv(W[:, :Cin], input) + conv(W[:, Cin:], hidden): the input half does not depend on the recurrent state and is
computed off the loop-carried critical path (on a side stream next to the hidden-state warp, or -- `input_gates()` --
in an earlier pipeline stage); only the hidden half (half of the 75 MB of weights) remains between two keyframes."""
import torch
import torch.nn as nn

from . import _native as N
from . import _ops as ops
from ._base import NativeModule

import os as _os
_FUSED_GATES = _os.environ.get("DVMVS_FUSED_GATES", "1") == "1"      # gate epilogue as the finishing pass of the gate GEMM


class MVSLayernormConvLSTMCell(NativeModule):
    def __init__(self, input_dim, hidden_dim, kernel_size, activation_function=None):
        super().__init__()
        self.activation_function = activation_function     # kept for signature parity; the kernel implements CELU(alpha=1)
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.kernel_size = kernel_size
        self.padding = kernel_size[0] // 2, kernel_size[1] // 2
        if tuple(kernel_size) != (3, 3) and tuple(kernel_size) != (1, 1) and tuple(kernel_size) != (5, 5):
            raise ValueError("kernel_size must be (1,1), (3,3) or (5,5)")
        if activation_function is not None and activation_function is not torch.celu:
            raise NotImplementedError("the fused gate kernel implements torch.celu (the reference's choice, fusionnet/model.py:319)")
        self.conv = nn.Conv2d(in_channels=input_dim + hidden_dim, out_channels=4 * hidden_dim, kernel_size=self.kernel_size,
                              padding=self.padding, bias=False)

    def _pack(self):
        w = self.conv.weight
        conv_x = ops.ConvLayer(ops.PackedConv(w[:, :self.input_dim], None, None, stride=1, act=N.ACT_NONE))
        conv_h = ops.ConvLayer(ops.PackedConv(w[:, self.input_dim:], None, None, stride=1, act=N.ACT_NONE))
        return conv_x, conv_h

    @ops.family_terms("lstm")
    def input_gates(self, input_tensor):
        """The state-independent half of the gate pre-activations, conv(W[:, :input_dim], input): (B,4*hidden,h,w).  Pass it
        to forward(..., input_gates=) to keep it off the recurrent critical path (PipelinedFusionnet does)."""
        conv_x, _ = self.packed()
        g = conv_x.run([(ops.to_act(input_tensor, "input_tensor"), N.SRC_DIRECT)], want_planes=False)
        return ops.to_api(g.f32)

    def forward(self, input_tensor, cur_state, previous_pose, current_pose, estimated_current_depth, camera_matrix, input_gates=None):
        conv_x, conv_h = self.packed()
        h_cur, c_cur = cur_state
        h = ops.to_act(h_cur, "h_cur")
        c = ops.to_nhwc(c_cur, "c_cur")
        fork = None
        if input_gates is None:
            fork = ops.Fork()                       # input half on a side stream, next to the hidden-state warp
            with fork:
                gx = conv_x.run([(ops.to_act(input_tensor, "input_tensor"), N.SRC_DIRECT)], want_planes=False)
        else:
            gx = ops.Act(ops.to_nhwc(input_gates, "input_gates"))
        if previous_pose is not None:
            # convlstm.py:30-41: transformation = inverse(previous_pose) @ current_pose, warp, mask depth <= 0.01
            h = ops.Act(ops.hidden_warp(h.f32, estimated_current_depth, previous_pose, current_pose, camera_matrix, 0.01))
        if fork is not None:
            fork.join()
        # the finishing pass of the (split-K) gate GEMM IS the gate epilogue: partial sums + input half -> sigmoid / LN / CELU / state
        deferred = conv_h.run_deferred([(h, N.SRC_DIRECT)]) if _FUSED_GATES else None
        if deferred is not None:
            h_next, c_next = ops.lstm_gates(None, c, parts=deferred[1], addend=gx.f32)
        else:
            gates = conv_h.run([(h, N.SRC_DIRECT)], residual=gx, residual_mode=N.RES_SAME, want_planes=False)
            h_next, c_next = ops.lstm_gates(gates.f32, c)
        return ops.act_to_api(ops.Act(h_next)), ops.to_api(c_next)

    def init_hidden(self, batch_size, image_size):
        height, width = image_size
        dev = self.conv.weight.device
        z = torch.zeros(batch_size, height, width, self.hidden_dim, device=dev)
        return ops.to_api(z), ops.to_api(torch.zeros_like(z))
