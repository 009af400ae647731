"""Oracle: functional CPU restatement of HardNet.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Parity: PINNED against
``/root/reference/third_party/hardnet/hardnet_pytorch.py`` (``tests/golden/hardnet_*.npz``).

Reference followed: hardnet_pytorch.py:36-59 (conv stack), :62-67 (input_norm),
:69-72 (forward), :7-15 (L2Norm).
"""
import torch
import torch.nn.functional as F

# (features index of conv, index of its BatchNorm, stride, padding, relu after)
LAYERS = ((0, 1, 1, 1, True), (3, 4, 1, 1, True), (6, 7, 2, 1, True), (9, 10, 1, 1, True),
          (12, 13, 2, 1, True), (15, 16, 1, 1, True), (19, 20, 1, 0, False))


def input_norm(x):
    """per-patch (x - mean) / (unbiased std + 1e-7)   (hardnet_pytorch.py:62-67)."""
    flat = x.reshape(x.shape[0], -1)
    mean = flat.mean(1).view(-1, 1, 1, 1)
    std = flat.std(1).view(-1, 1, 1, 1) + 1e-7
    return (x - mean) / std


def hardnet_forward(sd, patches):
    """patches [N,1,32,32] -> [N,128] (eval: BN running stats, affine=False; Dropout no-op)."""
    x = input_norm(patches)
    for conv, bn, stride, pad, relu in LAYERS:
        x = F.conv2d(x, sd["features.%d.weight" % conv], None, stride, pad)
        x = F.batch_norm(x, sd["features.%d.running_mean" % bn], sd["features.%d.running_var" % bn],
                         None, None, False, 0.0, 1e-5)
        if relu:
            x = F.relu(x)
    x = x.reshape(x.shape[0], -1)
    return x / torch.sqrt((x * x).sum(1) + 1e-10).unsqueeze(-1)
