"""COCO keypoint AP for HuPR without pycocotools: mirror of the reference's vendored evaluator for the case HuPR produces.

Follows /root/reference/misc/cocoeval.py (``COCOeval`` with ``iouType='keypoints'``: ``computeOks`` :192-236, ``evaluateImg``
:238-314, ``accumulate`` :316-420, ``_summarizeKps`` :475-487, ``Params.setKpParams`` :514-528) and the ``loadRes`` keypoint branch of
/root/reference/misc/coco.py:352-361, as driven by ``HuPR3D_horivert.evaluate`` / ``evaluateEach``
(/root/reference/datasets/dataset.py:48-88).  HuPR's files have exactly ONE ground-truth pose per image (datasets/base.py:55-75) and
at most ONE predicted pose per image, all with score 1.0 (tools/base.py:124-147); that restriction is asserted, which turns the
per-image greedy matching into one comparison per IoU threshold.  Everything else (tie order of equal scores, ignore rules of the
area ranges, the 101-point interpolated precision, the -1 conventions) is reproduced as written.

The similarity itself is computed on the GPU (``hupr_keypoint_oks``, float64) when the poses are CUDA tensors — the evaluation tail
then needs no per-batch device->host transfer of heatmaps or keypoints; with numpy inputs the same closed form is
evaluated on the host (used by the CPU tests that pin this module against the reference's evaluator).
"""
import json

import numpy as np

KPT_OKS_SIGMAS = np.array([1.07, .87, .89, 1.07, .87, .89, 1., 1., .79, .72, .62, .79, .72, .62]) / 10.0      # cocoeval.py:527
IOU_THRS = np.linspace(.5, 0.95, int(np.round((0.95 - .5) / .05)) + 1, endpoint=True)
REC_THRS = np.linspace(.0, 1.00, int(np.round((1.00 - .0) / .01)) + 1, endpoint=True)
AREA_RNG = [[0 ** 2, 1e5 ** 2], [32 ** 2, 96 ** 2], [96 ** 2, 1e5 ** 2]]
AREA_LBL = ["all", "medium", "large"]
MAX_DETS = 20


def oks(pred_xy, gt_xy, gt_area, sigmas=KPT_OKS_SIGMAS):
    """pred_xy, gt_xy ``[n, k, 2]``, gt_area ``[n]`` -> (oks ``[n]``, per-joint similarities ``[n, k]``) as float64 numpy arrays.
    CUDA tensors run ``hupr_keypoint_oks``; numpy arrays take the host closed form."""
    try:
        import torch
    except ImportError:          # pragma: no cover
        torch = None
    if torch is not None and isinstance(pred_xy, torch.Tensor) and pred_xy.is_cuda:
        from .. import _C
        dev = pred_xy.device
        n, k = pred_xy.shape[0], pred_xy.shape[1]
        p = pred_xy.to(torch.float32).contiguous()
        g = torch.as_tensor(gt_xy).to(device=dev, dtype=torch.float32).contiguous()
        a = torch.as_tensor(np.asarray(gt_area, dtype=np.float64) if not isinstance(gt_area, torch.Tensor) else gt_area).to(device=dev, dtype=torch.float64).contiguous()
        sg = torch.as_tensor(np.asarray(sigmas, dtype=np.float64)).to(dev)
        out = torch.empty(n, dtype=torch.float64, device=dev)
        per = torch.empty((n, k), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _C.check(_C.lib().hupr_keypoint_oks(_C.ptr(p), _C.ptr(g), _C.ptr(a), _C.ptr(sg), n, k, _C.ptr(out), _C.ptr(per), _C.stream_ptr()),
                     "hupr_keypoint_oks")
        return out.cpu().numpy(), per.cpu().numpy()
    pred = np.asarray(pred_xy, dtype=np.float64)
    gt = np.asarray(gt_xy, dtype=np.float64)
    area = np.asarray(gt_area, dtype=np.float64)
    var = (np.asarray(sigmas, dtype=np.float64) * 2) ** 2
    d2 = ((pred - gt) ** 2).sum(axis=2)
    per = np.exp(-(d2 / var / (area[:, None] + np.spacing(1)) / 2))
    return per.sum(axis=1) / per.shape[1], per


def accumulate(sim, has_dt, gt_id, gt_area, dt_area):
    """Per-image similarities -> the 10 COCO keypoint statistics (``COCOeval.stats``).  Arrays are in ascending image-id order (the
    order ``evaluate`` walks ``p.imgIds``); ``has_dt[i]`` is False for an image without a prediction."""
    sim = np.asarray(sim, dtype=np.float64)
    has_dt = np.asarray(has_dt, dtype=bool)
    T, R, A = len(IOU_THRS), len(REC_THRS), len(AREA_RNG)
    precision = -np.ones((T, R, A))
    recall = -np.ones((T, A))
    thr = np.minimum(IOU_THRS, 1 - 1e-10)[:, None]                       # evaluateImg: iou = min([t, 1-1e-10])
    for a, (lo, hi) in enumerate(AREA_RNG):
        gt_ig = (gt_area < lo) | (gt_area > hi)                            # :254
        dt_out = (dt_area < lo) | (dt_area > hi)                           # :303
        matched = has_dt[None, :] & ~(sim[None, :] < thr)                  # :284 (`continue` unless ious >= iou)
        dtm = np.where(matched, gt_id[None, :], 0)                         # :294 dtMatches holds the GT id; 0 = unmatched
        dt_ig = np.where(matched, gt_ig[None, :], False) | ((dtm == 0) & dt_out[None, :])
        dtm, dt_ig = dtm[:, has_dt], dt_ig[:, has_dt]                      # detections in image order: equal scores, stable sort
        npig = np.count_nonzero(~gt_ig)
        if npig == 0:
            continue
        tps = np.logical_and(dtm, np.logical_not(dt_ig))
        fps = np.logical_and(np.logical_not(dtm), np.logical_not(dt_ig))
        tp_sum = np.cumsum(tps, axis=1).astype(np.float64)
        fp_sum = np.cumsum(fps, axis=1).astype(np.float64)
        for t in range(T):
            tp, fp = tp_sum[t], fp_sum[t]
            nd = len(tp)
            rc = tp / npig
            pr = tp / (fp + tp + np.spacing(1))
            recall[t, a] = rc[-1] if nd else 0
            pr = np.maximum.accumulate(pr[::-1])[::-1] if nd else pr       # :392-394 monotone envelope
            inds = np.searchsorted(rc, REC_THRS, side="left")
            q = np.zeros(R)
            ok = inds < nd                                                 # :397-402 (the bare except leaves the tail at zero)
            q[ok] = pr[inds[ok]]
            precision[t, :, a] = q

    def summarize(ap, iou_thr=None, area="all"):
        a = AREA_LBL.index(area)
        s = precision[:, :, a] if ap else recall[:, a]
        if iou_thr is not None:
            s = s[np.where(iou_thr == IOU_THRS)[0]]
        return -1 if len(s[s > -1]) == 0 else float(np.mean(s[s > -1]))
    return np.array([summarize(1), summarize(1, .5), summarize(1, .75), summarize(1, area="medium"), summarize(1, area="large"),
                     summarize(0), summarize(0, .5), summarize(0, .75), summarize(0, area="medium"), summarize(0, area="large")])


class KeypointEval(object):
    """``KeypointEval(gt).evaluate(results)`` == ``COCOeval(COCO(gt), COCO(gt).loadRes(results), 'keypoints')`` evaluate / accumulate /
    summarize -> ``stats``; ``evaluate(results, idx_keypoint=j)`` is the reference's per-joint variant (cocoeval.py:121, :232-233)."""

    def __init__(self, gt, sigmas=KPT_OKS_SIGMAS):
        if isinstance(gt, str):
            with open(gt) as fp:
                gt = json.load(fp)
        anns = sorted(gt["annotations"], key=lambda a: a["image_id"])
        ids = [a["image_id"] for a in anns]
        if len(set(ids)) != len(ids) or sorted(img["id"] for img in gt["images"]) != ids:
            raise ValueError("KeypointEval handles HuPR ground truth: exactly one annotation per image")
        for a in anns:
            if a.get("iscrowd", 0) or a.get("num_keypoints", 1) == 0 or not all(v > 0 for v in a["keypoints"][2::3]):
                raise ValueError("KeypointEval handles HuPR ground truth: no crowd / unlabelled joints")
        self.image_ids = np.asarray(ids)
        self.index = {i: n for n, i in enumerate(ids)}
        self.gt_id = np.asarray([a["id"] for a in anns])
        self.gt_area = np.asarray([a["area"] for a in anns], dtype=np.float64)
        kp = np.asarray([a["keypoints"] for a in anns], dtype=np.float64).reshape(len(anns), -1, 3)
        self.gt_xy = kp[:, :, :2].copy()
        self.sigmas = np.asarray(sigmas, dtype=np.float64)
        self.stats = None

    def _gather(self, results):
        if isinstance(results, str):
            with open(results) as fp:
                results = json.load(fp)
        n, k = self.gt_xy.shape[0], self.gt_xy.shape[1]
        pred = np.zeros((n, k, 2))
        has = np.zeros(n, dtype=bool)
        for r in results:
            i = self.index.get(r["image_id"])
            if i is None:
                raise ValueError("results refer to image %r, which the ground truth does not hold" % (r["image_id"],))
            if has[i] or r.get("score", 1.0) != 1.0:
                raise ValueError("KeypointEval handles HuPR results: one pose per image, score 1.0")
            pred[i] = np.asarray(r["keypoints"], dtype=np.float64).reshape(k, 3)[:, :2]
            has[i] = True
        return pred, has

    def evaluate_arrays(self, pred_xy, has_dt=None, idx_keypoint=-1):
        """pred_xy ``[n_images, k, 2]`` in ascending image-id order (numpy, or a CUDA tensor -> similarities on the GPU)."""
        n = self.gt_xy.shape[0]
        has = np.ones(n, dtype=bool) if has_dt is None else np.asarray(has_dt, dtype=bool)
        total, per = oks(pred_xy, self.gt_xy, self.gt_area, self.sigmas)
        sim = total if idx_keypoint == -1 else per[:, idx_keypoint]
        host = pred_xy.detach().cpu().numpy().astype(np.float64) if hasattr(pred_xy, "detach") else np.asarray(pred_xy, dtype=np.float64)
        dt_area = (host[:, :, 0].max(1) - host[:, :, 0].min(1)) * (host[:, :, 1].max(1) - host[:, :, 1].min(1))     # coco.py:358-359
        self.stats = accumulate(sim, has, self.gt_id, self.gt_area, dt_area)
        return self.stats

    def evaluate(self, results, idx_keypoint=-1):
        pred, has = self._gather(results)
        return self.evaluate_arrays(pred, has, idx_keypoint)
