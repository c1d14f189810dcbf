"""Import surface of the reference's dvmvs/pairnet/model.py (identical to fusionnet's minus LSTMFusion)."""
from .._blocks import (CostVolumeDecoder, CostVolumeEncoder, DecoderBlock, DownconvolutionLayer, EncoderBlock,  # noqa: F401
                       FeatureExtractor, FeatureShrinker, StandardLayer, UpconvolutionLayer, fpn_output_channels,
                       hyper_channels)
from ..config import Config  # noqa: F401
from ..layers import conv_layer, depth_layer_3x3  # noqa: F401
