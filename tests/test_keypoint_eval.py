"""hupr_b200.misc.keypoint_eval against the reference's vendored COCO evaluator.

Golden: tests/golden/cocoeval_reference.npz, minted by ``python -m oracle.make_golden cocoeval`` from the UNMODIFIED
/root/reference/misc/{coco,cocoeval}.py (COCOeval 'keypoints': evaluate / accumulate / summarize, overall and per joint) on the seeded
HuPR-shaped files of ``oracle.make_golden.synth_keypoint_eval_case``.  Bar: the 10 statistics to 1e-12 (float64 throughout; they are
means of ratios of small integers), OKS to 1e-12."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cocoeval_reference.npz")


def _case(seed):
    from oracle.make_golden import synth_keypoint_eval_case
    return synth_keypoint_eval_case(seed)


@pytest.mark.parametrize("seed", [0, 1])
def test_host_mirror_matches_reference_cocoeval(seed):
    from hupr_b200.misc.keypoint_eval import KeypointEval
    gold = np.load(GOLDEN)
    gt, results = _case(seed)
    ev = KeypointEval(gt)
    stats = ev.evaluate(results)
    assert np.abs(stats - gold["seed%d_stats" % seed]).max() < 1e-12
    for j in range(14):
        sj = ev.evaluate(results, idx_keypoint=j)
        assert np.abs(sj - gold["seed%d_per_joint_stats" % seed][j]).max() < 1e-12, j
    # the similarities themselves
    pred, has = ev._gather(results)
    from hupr_b200.misc.keypoint_eval import oks
    total, _ = oks(pred, ev.gt_xy, ev.gt_area)
    ref = gold["seed%d_oks" % seed]
    assert np.array_equal(np.isnan(ref), ~has)
    assert np.abs(total[has] - ref[has]).max() < 1e-12


def test_rejects_files_outside_the_hupr_shape():
    from hupr_b200.misc.keypoint_eval import KeypointEval
    gt, results = _case(0)
    two = dict(gt, annotations=gt["annotations"] + [dict(gt["annotations"][0], id=1)])
    with pytest.raises(ValueError):
        KeypointEval(two)
    ev = KeypointEval(gt)
    with pytest.raises(ValueError):
        ev.evaluate(results + [results[0]])                       # two poses for one image
    with pytest.raises(ValueError):
        ev.evaluate([dict(results[0], image_id=-5)])


def test_perfect_and_empty_predictions():
    from hupr_b200.misc.keypoint_eval import KeypointEval
    gt, _ = _case(1)
    ev = KeypointEval(gt)
    stats = ev.evaluate_arrays(ev.gt_xy.copy())
    assert stats[0] == pytest.approx(1.0) and stats[5] == pytest.approx(1.0)
    stats = ev.evaluate_arrays(ev.gt_xy.copy(), has_dt=np.zeros(len(ev.gt_xy), dtype=bool))
    assert stats[0] == 0.0 and stats[5] == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1])
def test_device_oks_matches_reference_cocoeval(seed):
    """hupr_keypoint_oks (float64 on the GPU) -> identical statistics and similarities."""
    import torch
    from hupr_b200.misc.keypoint_eval import KeypointEval, oks
    gold = np.load(GOLDEN)
    gt, results = _case(seed)
    ev = KeypointEval(gt)
    pred, has = ev._gather(results)
    dev_pred = torch.from_numpy(pred).float().cuda()              # the results hold float32 values, so this is exact
    stats = ev.evaluate_arrays(dev_pred, has)
    assert np.abs(stats - gold["seed%d_stats" % seed]).max() < 1e-12
    sj = ev.evaluate_arrays(dev_pred, has, idx_keypoint=6)
    assert np.abs(sj - gold["seed%d_per_joint_stats" % seed][6]).max() < 1e-12
    total, per = oks(dev_pred, ev.gt_xy, ev.gt_area)
    ref = gold["seed%d_oks" % seed]
    assert np.abs(total[has] - ref[has]).max() < 1e-12
    host_total, host_per = oks(pred, ev.gt_xy, ev.gt_area)
    assert np.abs(per - host_per).max() < 1e-12
