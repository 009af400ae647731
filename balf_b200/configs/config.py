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


# multi-scale extraction: the defaults of the reference's parser (balf/configs/config_hpatches.py:50-82, parse_multiscale_config;
# the script that would consume them is not shipped)
MULTISCALE_DEFAULTS = dict(nms_size=15, num_points=1500, border_size=15, scale_factor_levels=2 ** 0.5, pyramid_levels=5,
                           upsampled_levels=1)


def default_multiscale_args(**overrides):
    return types.SimpleNamespace(**{**MULTISCALE_DEFAULTS, **overrides})


def multiscale_pyramid(margs):
    """parser arguments -> keyword arguments of demo_match.detect_multiscale_batch_device: the original image, `pyramid_levels`
    levels each 1 / scale_factor_levels of the previous one, and `upsampled_levels` levels each scale_factor_levels larger."""
    return dict(scale=1.0 / float(margs.scale_factor_levels), levels=int(margs.pyramid_levels) + 1,
                upsampled_levels=int(margs.upsampled_levels))
