"""Drop-in for the reference's ``balf/model/decoder.py`` (DetectorHead, decoder.py:5-30).

Parameters only: ``dense`` (Linear C -> cell^2+1) and ``norm`` (BatchNorm2d, eval statistics).
ReLU -> dense -> BN -> softmax -> drop dustbin -> depth-to-space is fused into the last
detector kernel (balf_b200/csrc/detector.cu); there is no eager path.
"""
import torch.nn as nn


class DetectorHead(nn.Module):
    def __init__(self, input_channel, cell_size):
        super().__init__()
        self.cell_size = cell_size
        self.dense = nn.Linear(input_channel, cell_size * cell_size + 1)
        self.norm = nn.BatchNorm2d(cell_size * cell_size + 1)

    def forward(self, x):
        raise RuntimeError("DetectorHead runs inside balf_detector_forward (CUDA); call MLP_MA_DECODER.forward")
