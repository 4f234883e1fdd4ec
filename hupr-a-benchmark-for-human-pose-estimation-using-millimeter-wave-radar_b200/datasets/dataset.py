"""HuPR window loader: host-side mirror of /root/reference/datasets/dataset.py (``getDataset`` / ``HuPR3D_horivert``) and
datasets/base.py (``generateGTAnnot``).

Same constructor, item dictionary (``VRDAEmap_hori``, ``VRDAEmap_vert``, ``imageId``, ``jointsGroup``, ``bbox``) and on-disk layout
(``<dataDir>/single_<g>/{hori,vert}/%09d.npy`` cubes ``[16,64,64,8]`` complex, ``<dataDir>/hrnet_annot_<phase>.json``,
generated ``<dataDir>/<phase>_gt.json`` in COCO keypoint format).  What changes is WHERE the loader arithmetic runs: the reference
normalises 256 planes per sample on the CPU (dataset.py:139-150, base.py:13-24; 1.25 s/sample); here the cube slices are uploaded
once per frame and ``hupr_window_normalize`` standardises them on the GPU, with a small per-frame cache because consecutive windows
share 7 of their 8 frames.  Items therefore carry CUDA tensors and the DataLoader must run with ``num_workers=0``.
COCO evaluation (``evaluate`` / ``evaluateEach``) uses ``hupr_b200.misc.keypoint_eval``, a mirror of the reference's vendored evaluator
(misc/cocoeval.py) pinned to it by golden statistics; ``evaluatePoses`` evaluates device-resident poses without the results file.
"""
import collections
import json
import os
import random as _random

import numpy as np
import torch
import torch.utils.data as data

from .. import ops


def window_frame_indices(index, duration, group=8):
    """dataset.py:126-138 in closed form: slot j reads frame clamp(index - group/2 + j) within the capture of ``duration`` frames."""
    first = index - index % duration
    last = first + duration - 1
    return [min(max(index - group // 2 + j, first), last) for j in range(group)]


def generateGTAnnot(cfg, phase="train"):
    """Write ``<dataDir>/<phase>_gt.json`` (COCO keypoint format) from ``hrnet_annot_<phase>.json`` — base.py:26-92.
    image_id = frame number + 100000 * capture group; bbox converted from corner to (x, y, w, h) form; all joints visible."""
    groups = getattr(cfg.DATASET, phase + "Name")
    annot = {"info": {"description": "HuPR dataset", "url": "", "version": "1.0", "year": 2022, "contributor": "UW-NYCU-AI-Labs",
                      "date_created": "2022/06/23"},
             "licenses": [], "images": [], "annotations": [],
             "categories": [{"supercategory": "person", "id": 1, "name": "person", "keypoints": list(cfg.DATASET.idxToJoints),
                             "skeleton": [[14, 13], [13, 12], [11, 10], [10, 9], [9, 7], [12, 9], [8, 7], [7, 1], [7, 4], [6, 5], [5, 4],
                                          [3, 2], [2, 1]]}]}
    with open(os.path.join(cfg.DATASET.dataDir, "hrnet_annot_%s.json" % phase)) as fp:
        blocks = json.load(fp)
    for i, sequence in enumerate(blocks):
        for item in sequence:
            image_id = int(item["image"][:-4]) + groups[i] * 100000
            x0, y0, x1, y1 = item["bbox"]
            keypoints = np.concatenate((np.asarray(item["joints"], dtype=np.float64), np.full((14, 1), 2.0)), axis=1).reshape(-1).tolist()
            annot["annotations"].append({"num_keypoints": 14, "area": (x1 - x0) * (y1 - y0) / 2, "iscrowd": 0, "keypoints": keypoints,
                                         "image_id": image_id, "bbox": [x0, y0, x1 - x0, y1 - y0], "category_id": 1, "id": image_id})
            annot["images"].append({"license": -1, "file_name": item["image"], "coco_url": "None", "height": 256, "width": 256,
                                    "date_captured": "None", "flickr_url": "None", "id": image_id})
    path = os.path.join(cfg.DATASET.dataDir, "%s_gt.json" % phase)
    with open(path, "w") as fp:
        json.dump(annot, fp)
    return path


def getDataset(phase, cfg, args, random=True):
    return HuPR3D_horivert(phase, cfg, args, random)


class HuPR3D_horivert(data.Dataset):
    def __init__(self, phase, cfg, args, random=True, device="cuda", cache_frames=64):
        if phase not in ("train", "val", "test"):
            raise ValueError("Invalid phase: {}".format(phase))
        super(HuPR3D_horivert, self).__init__()
        self.phase = phase
        self.duration = cfg.DATASET.duration
        self.numFrames = cfg.DATASET.numFrames
        self.numGroupFrames = cfg.DATASET.numGroupFrames
        self.numChirps = cfg.DATASET.numChirps
        self.numKeypoints = cfg.DATASET.numKeypoints
        self.sampling_ratio = args.sampling_ratio
        self.dirRoot = cfg.DATASET.dataDir
        self.idxToJoints = cfg.DATASET.idxToJoints
        self.random = random
        self.device = torch.device(device)
        if (self.numFrames, self.numGroupFrames, self.numChirps) != (8, 8, 16):
            raise ValueError("hupr_b200 loader kernels are specialised for 8 frames x 8 kept chirps of 16")
        self.gtFile = generateGTAnnot(cfg, phase)
        with open(self.gtFile) as fp:
            gt = json.load(fp)
        # COCO.getImgIds() with no filter returns list(self.imgs.keys()): the FILE's insertion order (not sorted — train/val/test name lists are
        # not ascending), so dataset index i addresses the same sample as in the reference (dataset.py:38; sampling_ratio, seeded shuffles)
        self.imageIds = list(dict.fromkeys(img["id"] for img in gt["images"]))
        by_image = collections.defaultdict(list)
        for ann in gt["annotations"]:
            if not ann.get("iscrowd", 0):
                by_image[ann["image_id"]].append(ann)
        self.VRDAEPaths_hori, self.VRDAEPaths_vert, self.annots = [], [], []
        for image_id in self.imageIds:
            text = "%09d" % image_id
            group, frame = int(text[:4]), int(text[-4:])
            self.VRDAEPaths_hori.append(os.path.join(self.dirRoot, "single_%d/hori/%09d.npy" % (group, frame)))
            self.VRDAEPaths_vert.append(os.path.join(self.dirRoot, "single_%d/vert/%09d.npy" % (group, frame)))
            for ann in by_image[image_id]:
                kp = np.asarray(ann["keypoints"], dtype=np.float64).reshape(self.numKeypoints, 3)
                self.annots.append({"joints": kp[:, :2].copy(), "joints_vis": np.minimum(kp[:, 2:3], 1.0).repeat(2, axis=1),
                                    "bbox": ann["bbox"], "imageId": ann["image_id"]})
        self._cache = collections.OrderedDict()          # path -> float32 CUDA tensor [8, 2, 64, 64, 8] (standardised kept chirps)
        self._cache_frames = cache_frames
        self._coco = None

    # ------------------------------------------------------------------------------------------------ items
    def _frame_planes(self, path):
        """Standardised planes of one stored cube: ``[8 kept chirps, 2 (re, im), 64, 64, 8]`` float32 on the GPU."""
        hit = self._cache.get(path)
        if hit is not None:
            self._cache.move_to_end(path)
            return hit
        cube = np.load(path)
        if cube.shape == (8, 2, 64, 64, 8) and cube.dtype == np.float32:
            # compact cache (RadarObject(cacheFormat='planes') / datasets.cubecache.convert): already standardised
            planes = torch.from_numpy(cube).to(self.device)
        else:
            if cube.shape != (16, 64, 64, 8):
                raise ValueError("%s: expected a [16,64,64,8] cube or a float32 [8,2,64,64,8] plane cache, got %s %s" % (path, cube.dtype, (cube.shape,)))
            dev_cube = torch.from_numpy(np.ascontiguousarray(cube.astype(np.complex64))).to(self.device).unsqueeze(0)
            planes = ops.window_normalize(dev_cube, torch.zeros(1, dtype=torch.int32, device=self.device))[0]
        self._cache[path] = planes
        if len(self._cache) > self._cache_frames:
            self._cache.popitem(last=False)
        return planes

    def __getitem__(self, index):
        if self.random:
            index = index * _random.randint(1, self.sampling_ratio)
        else:
            index = index * self.sampling_ratio
        frames = window_frame_indices(index, self.duration, self.numGroupFrames)
        hori = torch.stack([self._frame_planes(self.VRDAEPaths_hori[i]) for i in frames])
        vert = torch.stack([self._frame_planes(self.VRDAEPaths_vert[i]) for i in frames])
        annot = self.annots[index]
        return {"VRDAEmap_hori": hori, "VRDAEmap_vert": vert, "imageId": annot["imageId"],
                "jointsGroup": torch.LongTensor(annot["joints"]), "bbox": torch.FloatTensor(annot["bbox"])}

    def __len__(self):
        return len(self.VRDAEPaths_hori) // self.sampling_ratio

    # ------------------------------------------------------------------------------------------------ COCO keypoint evaluation
    # The reference patches pycocotools with its own misc/coco.py + misc/cocoeval.py (README 'Evaluation'); hupr_b200.misc.keypoint_eval
    # mirrors that evaluator for HuPR's one-pose-per-image files (pinned to it by tests/golden/cocoeval_reference.npz), so no
    # pycocotools install is needed and the similarities can be computed on the GPU.
    def _keval(self):
        from ..misc.keypoint_eval import KeypointEval
        if self._coco is None:
            self._coco = KeypointEval(self.gtFile)
        return self._coco

    @staticmethod
    def _print_stats(stats):
        labels = ["AP", "AP .5", "AP .75", "AP (M)", "AP (L)", "AR", "AR .5", "AR .75", "AR (M)", "AR (L)"]
        print(" | ".join("%s %.3f" % (l, v) for l, v in zip(labels, stats)))

    def evaluate(self, loadDir):
        """dataset.py:48-66: AP of ``<loadDir>/<phase>_results.json`` against the generated ground-truth file."""
        stats = self._keval().evaluate(os.path.join(loadDir, "%s_results.json" % self.phase))
        self._print_stats(stats)
        return stats[0]

    def evaluateEach(self, loadDir):
        """dataset.py:68-88: the same with the similarity restricted to one joint at a time; returns the last joint's AP like the
        reference does."""
        res_file = os.path.join(loadDir, "%s_results.json" % self.phase)
        per_joint = [self._keval().evaluate(res_file, idx_keypoint=i)[0] for i in range(self.numKeypoints)]
        for name, ap in zip(self.idxToJoints, per_joint):
            print("%s: %.3f" % (name, ap))
        return per_joint[-1]

    def evaluatePoses(self, imageIds, pred_xy, idx_keypoint=-1):
        """Evaluation tail without the results file: ``pred_xy`` ``[n, 14, 2]`` image-pixel poses (a CUDA tensor
        stays on the device: hupr_keypoint_oks) for the images ``imageIds``; images of the split without a pose count as missed."""
        ev = self._keval()
        order = [ev.index[int(i)] for i in imageIds]
        if len(set(order)) != len(order):
            raise ValueError("evaluatePoses: one pose per image")
        n = len(ev.image_ids)
        has = np.zeros(n, dtype=bool)
        has[order] = True
        if isinstance(pred_xy, torch.Tensor):
            full = torch.zeros((n,) + tuple(pred_xy.shape[1:]), dtype=torch.float32, device=pred_xy.device)
            full[torch.as_tensor(order, device=pred_xy.device)] = pred_xy.float()
        else:
            full = np.zeros((n,) + tuple(np.shape(pred_xy)[1:]))
            full[order] = np.asarray(pred_xy)
        return ev.evaluate_arrays(full, has, idx_keypoint)
