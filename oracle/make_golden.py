"""Generate tests/golden/* by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden [cascade|dca|loader|model|loss|cocoeval|all]

The fixtures pin the oracle restatements (oracle/*.py) to the reference's own code; they are small
(strided samples + float64 checksums) so they can live in git.  /root/reference is needed to run this
script, never to run the tests.
"""
import os
import sys
import tempfile

import numpy as np

from . import cascade, ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASCADE_CASES = [(0, 0), (3, 1)]           # (frame_idx, sensor) seeds of cascade.synth_frame
SAMPLE = (slice(None, None, 2), slice(None, None, 8), slice(None, None, 8), slice(None))


def cube_checksums(cube):
    cube = np.asarray(cube, dtype=np.complex128)
    return np.array([cube.real.sum(), cube.imag.sum(), (np.abs(cube) ** 2).sum(), np.abs(cube).max()],
                    dtype=np.float64)


def make_cascade():
    RadarObject = ref_shim.load_radar_object()
    ro = RadarObject()
    out = {}
    for frame_idx, sensor in CASCADE_CASES:
        frame = cascade.synth_frame(frame_idx, sensor)
        ref = ro.generateHeatmap(frame)
        key = "f%d_s%d" % (frame_idx, sensor)
        out[key + "_sample"] = ref[SAMPLE]
        out[key + "_checksums"] = cube_checksums(ref)
        out[key + "_dopplerplane_absmean"] = np.abs(ref).mean(axis=(1, 2, 3))
        print("cascade", key, ref.shape, out[key + "_checksums"])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "cascade_reference.npz"), **out)


def make_dca():
    RadarObject = ref_shim.load_radar_object()
    ro = RadarObject()
    rng = np.random.default_rng(7)
    n_frames = 2
    words = rng.integers(-32768, 32768, n_frames * cascade.FRAME_I16).astype(np.int16)
    with tempfile.TemporaryDirectory() as d:
        words.tofile(os.path.join(d, "adc_data.bin"))
        ref = ro.getadcDataFromDCA1000(d)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "dca1000_reference.npz"),
                        seed=np.array(7), n_frames=np.array(n_frames), shape=np.array(ref.shape),
                        sample=ref[:, ::37, ::5], real_sum=np.array(ref.real.sum()), imag_sum=np.array(ref.imag.sum()),
                        weighted=np.array((ref * np.arange(ref.size).reshape(ref.shape)).sum()))
    print("dca", ref.shape)


MODEL_SAMPLE = (slice(None), slice(None), slice(None), slice(3, None, 7), slice(2, None, 5))


def tensor_stats(t):
    t = t.double()
    return np.array([float(t.sum()), float((t * t).sum()), float(t.abs().max())], dtype=np.float64)


def make_model():
    """Reference HuPRNet (eval) on the oracle's seeded weights/inputs -> strided samples + checksums."""
    import torch
    from . import model as om
    cls = ref_shim.load_model_classes()
    net = cls["HuPRNet"](ref_shim.load_cfg()).eval()
    out = {}
    for batch, seed in ((1, 0), (2, 5)):
        sd = om.make_state_dict(seed)
        assert list(net.state_dict().keys()) == list(sd.keys())
        net.load_state_dict(sd)
        hori, vert = om.make_vrdae(batch, seed)
        with torch.no_grad():
            heat, gcn = net(hori, vert)
            ra = net.forward_chirp(hori, vert)[0]
            feats = net.RAradarEncoder(ra)
        key = "b%d_s%d" % (batch, seed)
        out[key + "_heatmap"] = heat.numpy()[MODEL_SAMPLE]
        out[key + "_gcn"] = gcn.numpy()[MODEL_SAMPLE]
        out[key + "_heatmap_stats"] = tensor_stats(heat)
        out[key + "_gcn_stats"] = tensor_stats(gcn)
        out[key + "_chirp_stats"] = tensor_stats(ra)
        for i, f in enumerate(feats):
            out[key + "_enc%d_stats" % i] = tensor_stats(f)
            out[key + "_enc%d" % i] = f.numpy()[:, ::8, ::5, ::5]
        print("model", key, out[key + "_heatmap_stats"], out[key + "_gcn_stats"])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "model_reference.npz"), **out)


def make_loss():
    """Reference LossComputer / generateTarget / get_max_preds on seeded predictions and joints."""
    import torch
    LossComputer, generateTarget, get_max_preds = ref_shim.load_misc()
    lc = LossComputer(ref_shim.load_cfg(), torch.device("cpu"))
    g = torch.Generator().manual_seed(42)
    b = 3
    heat = torch.rand((b, 14, 1, 64, 64), generator=g) * 0.98 + 0.01
    gcn = torch.rand((b, 1, 14, 64, 64), generator=g) * 0.98 + 0.01
    # joints include border / out-of-range cases so that the Gaussian clipping branches are covered
    gt = torch.randint(0, 256, (b, 14, 2), generator=g)
    gt[0, 0] = torch.tensor([0, 0]); gt[0, 1] = torch.tensor([255, 255]); gt[0, 2] = torch.tensor([3, 250])
    gt[1, 0] = torch.tensor([300, 10]); gt[1, 1] = torch.tensor([-40, 100]); gt[1, 2] = torch.tensor([283, 128])
    loss, loss2, pred2d, gt2d = lc.computeLoss((heat, gcn), gt)
    targets = np.stack([generateTarget(gt[i], 14, 64, 256)[0] for i in range(b)])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "loss_reference.npz"), seed=np.array(42), gt=gt.numpy(),
                        loss=np.array(float(loss)), loss2=np.array(float(loss2)), pred2d=pred2d, gt2d=gt2d,
                        targets_sample=targets[:, :, ::3, ::3], targets_sum=np.array(targets.astype(np.float64).sum()),
                        targets_nonzero=np.array(int((targets > 0).sum())))
    print("loss", float(loss), float(loss2))


def make_loader():
    """Reference Normalize (through torchvision ToTensor, as getTransformFunc builds it) on cascade cubes."""
    import torchvision.transforms as transforms
    Normalize = ref_shim.load_normalize()
    tf = transforms.Compose([transforms.ToTensor(), Normalize()])
    cube = cascade.generate_heatmap(cascade.synth_frame(2, 0))
    out = np.zeros((8, 2, 64, 64, 8), dtype=np.float32)
    for s, d in enumerate(range(4, 12)):
        out[s, 0] = tf(cube[d].real).permute(1, 2, 0).numpy()
        out[s, 1] = tf(cube[d].imag).permute(1, 2, 0).numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "loader_reference.npz"), frame=np.array([2, 0]),
                        sample=out[:, :, ::4, ::4, :], sums=out.astype(np.float64).sum(axis=(2, 3)),
                        sqsums=(out.astype(np.float64) ** 2).sum(axis=(2, 3)))
    print("loader", out.shape, np.isfinite(out).all())



def synth_keypoint_eval_case(seed, n_images=60, k=14):
    """Seeded COCO-format ground truth + results in the shape HuPR writes them (datasets/base.py:26-92, tools/base.py:124-147): one
    pose per image, integer joints in a 256x256 frame, all joints visible, area = bbox area / 2; results = the ground truth displaced by
    noise whose scale varies per image (so the similarities straddle every IoU threshold), boxes spanning all three area ranges, a few
    images without a prediction."""
    rng = np.random.default_rng(seed)
    gt = {"images": [], "annotations": [], "categories": [{"supercategory": "person", "id": 1, "name": "person",
                                                           "keypoints": ["j%d" % j for j in range(k)], "skeleton": []}]}
    results = []
    for i in range(n_images):
        image_id = 100000 * (1 + i // 25) + (i % 25)
        size = float(rng.choice([20, 40, 70, 110, 160]))                    # bbox side: areas/2 of 200 .. 12 800 px^2
        x0, y0 = rng.uniform(0, 256 - size, 2)
        joints = np.rint(np.stack([rng.uniform(x0, x0 + size, k), rng.uniform(y0, y0 + size, k)], axis=1))
        kp = np.concatenate([joints, np.full((k, 1), 2.0)], axis=1).reshape(-1).tolist()
        gt["images"].append({"id": image_id, "height": 256, "width": 256, "file_name": "%09d.jpg" % image_id})
        gt["annotations"].append({"num_keypoints": k, "area": size * size / 2, "iscrowd": 0, "keypoints": kp, "image_id": image_id,
                                  "bbox": [float(x0), float(y0), size, size], "category_id": 1, "id": image_id})
        if rng.uniform() < 0.08:
            continue                                                          # no prediction for this image
        noise = rng.choice([0.0, 1.0, 3.0, 6.0, 12.0]) * size / 64.0
        pred = np.float32(joints + rng.normal(0, 1, (k, 2)) * noise)          # float32 like pred2d * imgHeatmapRatio
        res_kp = np.concatenate([pred.astype(np.float64), np.ones((k, 1))], axis=1).reshape(-1).tolist()
        results.append({"category_id": 1, "image_id": image_id, "score": 1.0, "keypoints": res_kp})
    return gt, results


def make_cocoeval():
    """Reference vendored COCOeval (misc/coco.py, misc/cocoeval.py) on seeded HuPR-shaped files: the 10 summary statistics, overall
    and for every single joint (``evaluate(idx_keypoint)``, as datasets/dataset.py:68-88 drives it)."""
    import contextlib
    import io
    import json
    COCO, COCOeval = ref_shim.load_cocoeval()
    out = {}
    for seed in (0, 1):
        gt, results = synth_keypoint_eval_case(seed)
        with tempfile.TemporaryDirectory() as tmp:
            gt_path, res_path = os.path.join(tmp, "gt.json"), os.path.join(tmp, "res.json")
            with open(gt_path, "w") as fp:
                json.dump(gt, fp)
            with open(res_path, "w") as fp:
                json.dump(results, fp)
            with contextlib.redirect_stdout(io.StringIO()):
                coco = COCO(gt_path)
                ev = COCOeval(coco, coco.loadRes(res_path), "keypoints")
                ev.params.useSegm = None
                ev.evaluate()
                ev.accumulate()
                ev.summarize()
                stats = np.array(ev.stats)
                per_joint = []
                for j in range(14):
                    ev.evaluate(j)
                    ev.accumulate()
                    ev.summarize()
                    per_joint.append(np.array(ev.stats))
                oks_ref = np.array([float(ev.computeOks(a["image_id"], 1)[0][0]) if len(ev.computeOks(a["image_id"], 1)) else np.nan
                                    for a in sorted(gt["annotations"], key=lambda a: a["image_id"])])
        out["seed%d_stats" % seed] = stats
        out["seed%d_per_joint_stats" % seed] = np.stack(per_joint)
        out["seed%d_oks" % seed] = oks_ref
        print("cocoeval seed", seed, "stats", np.round(stats, 4))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "cocoeval_reference.npz"), **out)


TRAIN_CASE = dict(batch=2, seed=1, joints_seed=3)       # the inputs tests/test_training_step_gpu.py feeds the whole-step tests
GRAD_STRIDE = 499                                        # every 499th element of each flattened gradient is stored


def train_case_inputs():
    import torch
    from . import model as om
    sd = om.make_state_dict(TRAIN_CASE["seed"])
    hori, vert = om.make_vrdae(TRAIN_CASE["batch"], TRAIN_CASE["seed"])
    joints = torch.randint(0, 256, (TRAIN_CASE["batch"], 14, 2), generator=torch.Generator().manual_seed(TRAIN_CASE["joints_seed"]))
    return sd, hori, vert, joints


def make_train():
    """One reference training pass (tools/run.py:66,76-78): HuPRNet.train() forward (batch-statistics BatchNorm), LossComputer.computeLoss,
    loss.backward() — on the oracle's seeded weights / inputs.  Stored: the two losses, for each of the 165 parameters a strided
    sample + (sum, sum of squares, max |.|) of its gradient, and the running statistics the pass leaves in every BatchNorm."""
    import torch
    cls = ref_shim.load_model_classes()
    LossComputer, _, _ = ref_shim.load_misc()
    sd, hori, vert, joints = train_case_inputs()
    net = cls["HuPRNet"](ref_shim.load_cfg())
    net.load_state_dict(sd)
    net.train()
    lc = LossComputer(ref_shim.load_cfg(), torch.device("cpu"))
    preds = net(hori, vert)
    loss, loss2, pred2d, gt2d = lc.computeLoss(preds, joints)
    loss.backward()
    out = {"loss": np.array(float(loss.detach())), "loss2": np.array(float(loss2.detach())), "pred2d": np.asarray(pred2d), "gt2d": np.asarray(gt2d),
           "batch": np.array(TRAIN_CASE["batch"]), "seed": np.array(TRAIN_CASE["seed"]), "joints": joints.numpy(),
           "grad_stride": np.array(GRAD_STRIDE)}
    names = []
    for name, q in net.named_parameters():
        g = q.grad.detach().reshape(-1)
        names.append(name)
        out["g/" + name] = g[::GRAD_STRIDE].numpy().copy()
        out["s/" + name] = tensor_stats(g)
    for name, buf in net.named_buffers():
        if name.endswith("running_mean") or name.endswith("running_var"):
            out["b/" + name] = buf.detach().numpy().copy()
    out["names"] = np.array(names)
    assert len(names) == 165
    np.savez_compressed(os.path.join(GOLDEN_DIR, "train_reference.npz"), **out)
    print("train", float(loss), float(loss2), len(names), "gradients")


TARGETS = {"cascade": make_cascade, "dca": make_dca, "model": make_model, "loss": make_loss, "loader": make_loader,
           "cocoeval": make_cocoeval, "train": make_train}


def main(argv):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    which = argv[1:] or ["all"]
    for name, fn in TARGETS.items():
        if "all" in which or name in which:
            fn()


if __name__ == "__main__":
    main(sys.argv)
