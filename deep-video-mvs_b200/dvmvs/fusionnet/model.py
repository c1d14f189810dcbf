"""Import surface of the reference's dvmvs/fusionnet/model.py (run-testing.py:6, run-training.py:14 `import *`)."""
from .._blocks import (CostVolumeDecoder, CostVolumeEncoder, DecoderBlock, DownconvolutionLayer, EncoderBlock,  # noqa: F401
                       FeatureExtractor, FeatureShrinker, LSTMFusion, StandardLayer, UpconvolutionLayer,
                       fpn_output_channels, hyper_channels)
from ..config import Config  # noqa: F401
from ..convlstm import MVSLayernormConvLSTMCell  # noqa: F401
from ..layers import conv_layer, depth_layer_3x3  # noqa: F401
