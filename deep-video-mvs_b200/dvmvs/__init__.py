"""dvmvs -- B200-native drop-in for the plane-sweep depth-inference path of ardaduz/deep-video-mvs.

Same import surface as the reference package of the same name (dvmvs.fusionnet.model, dvmvs.pairnet.model,
dvmvs.utils, dvmvs.convlstm, dvmvs.layers, dvmvs.config, dvmvs.dataset_loader), so the reference's
fusionnet/run-testing.py and pairnet/run-testing.py run against it unchanged; every hot op is a hand-written
sm_100a CUDA kernel reached through the C ABI of libdvmvs_sm100.so (include/dvmvs_b200.h).  There is no CPU
or eager-PyTorch fallback: ops raise RuntimeError when the library is missing or a tensor is not on a GPU.
"""
__version__ = "0.1.0"
