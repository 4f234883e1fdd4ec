"""CPU check of the two-unit convolution ARITHMETIC (no kernel involved): the torch-CPU oracle network is re-run with the operands of its
3x3(x3) convolutions quantised exactly as hupr_conv_desc.nprod == 2 prescribes (include/hupr_b200.h: fp16 main product with scales 2^2 /
2^14, e4m3 cross terms with scales 2^12 x 2^4 and 2^1 x 2^15, saturating conversions) and every other contraction as three bf16
products; the heat maps must stay inside north_star's 1e-3 element-wise bar with margin and the argmax keypoints must not move.  This pins
the numbers DESIGN.md §2 quotes (tools_dev/precision_emulation.py, profiles/r02_precision_emulation.txt) in the CPU tier."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools_dev"))


def test_two_unit_arithmetic_keeps_the_oracle_heatmaps_within_tolerance():
    import precision_emulation as pe
    from oracle import loss as oloss
    from oracle import model as om
    saved = (pe.F.conv3d, pe.F.conv2d, om.torch.einsum, om.torch.matmul)
    try:
        pe.install()
        sd = om.make_state_dict(0)
        h, v = om.make_vrdae(1, 0)
        pe.MODE = None
        with torch.no_grad():
            rh, rg = om.huprnet_forward(sd, h, v)
        rk, _ = oloss.get_max_preds(rg.view(1, 14, 64, 64).numpy())
        errs = {}
        for mode in ("bf16x3", "halo-f16f8"):
            pe.MODE = mode
            with torch.no_grad():
                gh, gg = om.huprnet_forward(sd, h, v)
            errs[mode] = max(float(((gh - rh).abs() / rh.abs()).max()), float(((gg - rg).abs() / rg.abs()).max()))
            k, _ = oloss.get_max_preds(gg.view(1, 14, 64, 64).numpy())
            assert (k == rk).all(), mode
        print("element-wise relative heat-map error: three bf16 products %.3g, two-unit 3-tap convolutions %.3g" % (errs["bf16x3"], errs["halo-f16f8"]))
        assert errs["bf16x3"] < 1e-4                      # the emulation reproduces what the GPU path measures (2.7e-5 on these seeds)
        assert errs["halo-f16f8"] < 3e-4                  # VERDICT r1 item 4's bar for the cheaper correction products; measured 5.6e-5
    finally:
        pe.MODE = None
        pe.F.conv3d, pe.F.conv2d, om.torch.einsum, om.torch.matmul = saved
