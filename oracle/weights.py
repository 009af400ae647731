"""Oracle: seeded random-init weights of the detector and of HardNet, built WITHOUT importing the product.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  ``bench.py --impl reference`` and the CPU baseline use
these state_dicts so that the CPU arm never loads ``balf_b200`` (and with it ``libbalf_b200.so``).

The reference creates its parameters with the default ``nn.Linear`` / ``nn.LayerNorm`` / ``nn.BatchNorm2d`` /
``nn.Conv2d`` initialisers in constructor order, so replaying the same constructors in the same order under
``torch.manual_seed(seed)`` reproduces its weights bit for bit.  Order followed (reference file:line):
  Down.__init__                      balf/model/mlp_ma_decoder.py:202-221   conv.0, gMLP layer, channel attention, conv2
  ResidualSplitHeadMultiAxisGmlpLayer.__init__   :121-130   norm, dense1, grid layer, block layer, dense2
  GridGmlpLayer / BlockGmlpLayer.__init__        :46-55, :93-102   norm, dense1, gating unit (norm, dense), dense2
  ResidualChannelAttentionBlock.__init__         :175-183   norm, conv1, conv2, CALayer (excite.0, excite.2)
  DetectorHead.__init__              balf/model/decoder.py:6-12   dense, BatchNorm2d
  HardNet.__init__                   third_party/hardnet/hardnet_pytorch.py:32-58 (+ weights_init :74-81 is NOT applied
                                     by the constructor in this fork: default Conv2d init)
PINNED: tests/test_oracle_golden.py compares the digests with the reference's own (tests/golden/detector.npz, hardnet.npz).
"""
from collections import OrderedDict

import torch
import torch.nn as nn

DEFAULT_ARCH = dict(en_embed_dims=[3, 32, 64, 128, 256], grid_size=[8, 8], block_size=[8, 8], grid_gmlp_factor=2,
                    block_gmlp_factor=2, input_proj_factor=2, channels_reduction=4, cell_size=8)
_P = "residual_split_head_multi_axis_gmlp_layer"
_R = "residual_channel_attention_block"


def _emit(sd, name, mod):
    for k, v in mod.state_dict().items():
        sd[name + "." + k] = v.detach().clone()


def _gmlp(sd, pre, gate, c, tokens, factor):
    _emit(sd, pre + ".norm", nn.LayerNorm(c))
    _emit(sd, pre + ".dense1", nn.Linear(c, c * factor))
    _emit(sd, pre + "." + gate + ".norm", nn.LayerNorm(c))
    _emit(sd, pre + "." + gate + ".dense", nn.Linear(tokens, tokens))
    _emit(sd, pre + ".dense2", nn.Linear(c, c))


def detector_state_dict(seed=0, arch=None):
    a = dict(DEFAULT_ARCH if arch is None else arch)
    dims = a["en_embed_dims"]
    cells = a["grid_size"][0] * a["grid_size"][1]
    block_px = a["block_size"][0] * a["block_size"][1]
    torch.manual_seed(seed)
    sd = OrderedDict()
    for i in range(4):
        cin, c = dims[i], dims[i + 1]
        d = "down%d" % (i + 1)
        _emit(sd, d + ".conv.0", nn.Linear(cin, c))
        _emit(sd, d + "." + _P + ".norm", nn.LayerNorm(c))
        _emit(sd, d + "." + _P + ".dense1", nn.Linear(c, c * a["input_proj_factor"]))
        _gmlp(sd, d + "." + _P + ".grid_gmlp_layer", "grid_gating_unit", c, cells, a["grid_gmlp_factor"])
        _gmlp(sd, d + "." + _P + ".block_gmlp_layer", "block_gating_unit", c, block_px, a["block_gmlp_factor"])
        _emit(sd, d + "." + _P + ".dense2", nn.Linear(c * a["input_proj_factor"], c))
        _emit(sd, d + "." + _R + ".norm", nn.LayerNorm(c))
        _emit(sd, d + "." + _R + ".conv1", nn.Linear(c, c))
        _emit(sd, d + "." + _R + ".conv2", nn.Linear(c, c))
        _emit(sd, d + "." + _R + ".calayer.excite.0", nn.Linear(c, c // a["channels_reduction"]))
        _emit(sd, d + "." + _R + ".calayer.excite.2", nn.Linear(c // a["channels_reduction"], c))
        _emit(sd, d + ".conv2", nn.Linear(c, c))
    n = a["cell_size"] ** 2 + 1
    _emit(sd, "detector_head.dense", nn.Linear(dims[4], n))
    _emit(sd, "detector_head.norm", nn.BatchNorm2d(n))
    return sd


_HN = ((0, 1, 32, 3, 1, 1), (3, 32, 32, 3, 1, 1), (6, 32, 64, 3, 2, 1), (9, 64, 64, 3, 1, 1), (12, 64, 128, 3, 2, 1),
       (15, 128, 128, 3, 1, 1), (19, 128, 128, 8, 1, 0))


def hardnet_state_dict(seed=0):
    torch.manual_seed(seed)
    sd = OrderedDict()
    for idx, cin, cout, k, stride, pad in _HN:
        _emit(sd, "features.%d" % idx, nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=pad, bias=False))
        _emit(sd, "features.%d" % (idx + 1), nn.BatchNorm2d(cout, affine=False))
    return sd
