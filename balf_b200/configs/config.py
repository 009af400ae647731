"""Drop-in for ``balf/configs/config.py::parse_test_config`` (config.py:35-65): same flags and
defaults (border 15, nms 15, k 2048, s_mult 60, threshold 0.001, sub_pixel True, patch 4)."""
import argparse
import os
import types

from ..utils import test_utils

DEFAULT_CFG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "test.yaml")

TEST_DEFAULTS = dict(border_size=15, nms_size=15, num_features=2048, s_mult=60, order_coord="xysr",
                     heatmap_confidence_threshold=0.001, sub_pixel=True, patch_size=4)


def default_test_args(**overrides):
    """The argparse defaults as a namespace, for library use without a command line."""
    return types.SimpleNamespace(**{**TEST_DEFAULTS, **overrides})


def parse_test_config(argv=None):
    p = argparse.ArgumentParser(description="motion blur feature matching test")
    p.add_argument("--cfg_file", type=str, default=DEFAULT_CFG)
    p.add_argument("--ckpt_file", type=str, default="pretrained/balf/balf.pth")
    p.add_argument("--ckpt_descriptor_file", type=str, default="pretrained/hardnet/HardNet++.pth")
    p.add_argument("--border_size", type=int, default=15)
    p.add_argument("--nms_size", type=int, default=15)
    p.add_argument("--num_features", type=int, default=2048)
    p.add_argument("--s_mult", type=int, default=60)
    p.add_argument("--order_coord", type=str, default="xysr")
    p.add_argument("--heatmap_confidence_threshold", type=float, default=0.001)
    p.add_argument("--sub_pixel", type=bool, default=True)     # as in the reference: any non-empty string is True
    p.add_argument("--patch_size", type=int, default=4)
    args = p.parse_args(argv)
    return args, test_utils.get_cfg_from_yaml_file(args.cfg_file)
