"""Drop-in for the reference's ``balf/model/get_model.py``.

``load_model(model_cfg)`` (get_model.py:88-90) and ``load_test_pretrained_model(model, filename,
optimizer=None, device='cuda')`` (get_model.py:50-86) keep their signatures, return values and
error behaviour: FileNotFoundError for a missing file, "Not updated weight" print for keys the
checkpoint lacks, AssertionError unless every state_dict key was matched by name and shape.
"""
import os

import torch

from . import mlp_ma_decoder


def load_model(model_cfg):
    return mlp_ma_decoder.MLP_MA_DECODER(model_cfg["network_architecture"])


def _restore(model, filename, optimizer, device, log):
    if not os.path.isfile(filename):
        raise FileNotFoundError
    where = torch.device("cpu") if device == "cpu" else None
    ckpt = torch.load(filename, map_location=where, weights_only=False)
    own = model.state_dict()
    taken = {k: v for k, v in ckpt["model_state"].items() if k in own and own[k].shape == v.shape}
    own.update(taken)
    model.load_state_dict(own)
    for k in own:
        if k not in taken:
            log("Not updated weight %s: %s" % (k, str(own[k].shape)))
    if optimizer is not None:
        if ckpt.get("optimizer_state") is not None:
            optimizer.load_state_dict(ckpt["optimizer_state"])
        else:
            assert filename[-4] == ".", filename
            side = "%s_optim.%s" % (filename[:-4], filename[-3:])
            if os.path.exists(side):
                optimizer.load_state_dict(torch.load(side, map_location=where)["optimizer_state"])
    assert len(taken) == len(model.state_dict())
    return ckpt.get("epoch", -1), ckpt.get("repeatability", 0.0)


def load_test_pretrained_model(model, filename, optimizer=None, device="cuda"):
    return _restore(model, filename, optimizer, device, print)


def load_pretrained_model(model, filename, logger, optimizer=None, device="cuda"):
    """Training-side variant (get_model.py:6-48): same restore, messages go to ``logger``."""
    return _restore(model, filename, optimizer, device, logger.info)
