"""Train / evaluate driver: host-side mirror of /root/reference/tools/run.py (``Runner``) and tools/base.py (``BaseRunner``).

Same constructor ``Runner(args, cfg)``, methods (``loadModelWeight``, ``train``, ``eval``, ``saveModelWeight``, ``saveKeypoints``,
``writeKeypoints``, ``adjustLR``), directory conventions (``./logs/<dir>/{checkpoint,model_best}.pth``, ``<phase>_results.json``) and
checkpoint dictionary (``epoch``, ``model_state_dict``, ``optimizer_state_dict`` in torch.optim.Adam's own format, ``accuracy``), so
checkpoints written by either implementation load in the other.  The arithmetic runs through hupr_b200: ``HuPRNet`` (inference),
``TrainStep`` (forward + backward + Adam), ``LossComputer`` (loss / keypoint decode).
Differences: visualisation (``plotHumanPose``, needs RGB frames) is not provided; ``eval`` keeps the poses on the device, writes the
results file from one transfer after the loop and computes AP with ``hupr_b200.misc.keypoint_eval`` (a mirror of the reference's
vendored COCO evaluator), so no pycocotools install is needed.  One process drives one GPU; data-parallel training over several GPUs
is ``TrainStep`` + ``all_reduce_gradients`` under torchrun (``bench.py --workload train`` shows the loop).
"""
import json
import math
import os

import numpy as np
import torch
import torch.utils.data as data

from ..datasets import getDataset
from ..misc import LossComputer
from ..models import HuPRNet
from ..training import TrainStep


class Runner(object):
    def __init__(self, args, cfg):
        if not torch.cuda.is_available():
            raise RuntimeError("hupr_b200 needs a CUDA device (B200); there is no CPU path")
        self.device = "cuda"
        np.random.seed(args.seed)
        torch.manual_seed(args.seed)
        torch.cuda.manual_seed_all(args.seed)
        self.args, self.cfg = args, cfg
        self.dir = "./logs/" + args.dir
        self.visDir = "./visualization/" + args.visDir
        self.numKeypoints = cfg.DATASET.numKeypoints
        self.heatmapSize = cfg.DATASET.heatmapSize
        self.imgHeatmapRatio = cfg.DATASET.imgSize / cfg.DATASET.heatmapSize
        self.aspectRatio = 1.0
        self.pixel_std = 200
        self.start_epoch = 0
        self.bestAP = -1.0
        if not args.eval:
            self.trainSet = getDataset("train", cfg, args)
            self.trainLoader = data.DataLoader(self.trainSet, cfg.TRAINING.batchSize, shuffle=True, num_workers=0, drop_last=False)
        else:
            self.trainLoader = [0]
        self.testSet = getDataset("test" if args.eval else "val", cfg, args)
        self.testLoader = data.DataLoader(self.testSet, cfg.TEST.batchSize, shuffle=False, num_workers=0)
        self.model = HuPRNet(cfg).to(self.device)
        steps = len(self.trainLoader) * cfg.TRAINING.warmupEpoch
        lr = cfg.TRAINING.lr if cfg.TRAINING.warmupEpoch == -1 else cfg.TRAINING.lr / (cfg.TRAINING.warmupGrowth ** steps)
        self.initialize(lr)

    def initialize(self, LR):
        self.lossComputer = LossComputer(self.cfg, self.device)
        os.makedirs(self.dir, exist_ok=True)
        if self.cfg.TRAINING.optimizer != "adam":
            raise ValueError("hupr_b200 implements the reference's Adam configuration (TRAINING.optimizer: adam)")
        self.trainer = None if self.args.eval else TrainStep(self.model, lr=LR, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
        if not self.args.eval:
            print("==========>Train set size:", len(self.trainLoader))
        print("==========>Test set size:", len(self.testLoader))

    # ---------------------------------------------------------------------------------------------- schedule / checkpoints
    def adjustLR(self, epoch):
        factor = self.cfg.TRAINING.warmupGrowth if epoch < self.cfg.TRAINING.warmupEpoch else self.cfg.TRAINING.lrDecay
        self.trainer.hyper["lr"] *= factor

    def _optimizer_state_dict(self):
        """torch.optim.Adam.state_dict() layout, one entry per parameter in model.parameters() order."""
        t, state = self.trainer, {}
        for i, (name, p) in enumerate(self.model.named_parameters()):
            o, k = t.offsets[name], p.numel()
            state[i] = {"step": torch.tensor(float(t.step_count)), "exp_avg": t.exp_avg[o:o + k].view(p.shape).clone(),
                        "exp_avg_sq": t.exp_avg_sq[o:o + k].view(p.shape).clone()}
        group = dict(lr=t.hyper["lr"], betas=t.hyper["betas"], eps=t.hyper["eps"], weight_decay=t.hyper["weight_decay"], amsgrad=False,
                     params=list(range(len(state))))
        return {"state": state, "param_groups": [group]}

    def _load_optimizer_state_dict(self, sd):
        t = self.trainer
        for i, (name, p) in enumerate(self.model.named_parameters()):
            o, k = t.offsets[name], p.numel()
            entry = sd["state"].get(i)
            if entry is not None:
                t.exp_avg[o:o + k].copy_(entry["exp_avg"].reshape(-1))
                t.exp_avg_sq[o:o + k].copy_(entry["exp_avg_sq"].reshape(-1))
                t.step_count = int(float(entry["step"]))
        t.step_dev.fill_(t.step_count)
        t.set_lr(sd["param_groups"][0]["lr"])

    def saveModelWeight(self, epoch, acc):
        is_best = (acc == acc) and acc > self.bestAP
        if is_best:
            self.bestAP = acc
        group = {"epoch": epoch, "model_state_dict": self.model.state_dict(), "optimizer_state_dict": self._optimizer_state_dict(),
                 "accuracy": self.bestAP}
        if is_best:
            print("==========>Save the best model...")
            torch.save(group, os.path.join(self.dir, "model_best.pth"))
        print("==========>Save the latest model...")
        torch.save(group, os.path.join(self.dir, "checkpoint.pth"))
        if epoch % 5 == 0:
            torch.save(group, os.path.join(self.dir, "checkpoint_%d.pth" % epoch))

    def saveLosslist(self, epoch, loss_list, mode):
        with open(os.path.join(self.dir, "%s_loss_list_%d.json" % (mode, epoch)), "w") as fp:
            json.dump(loss_list, fp)

    def loadModelWeight(self, mode):
        path = os.path.join(self.dir, "%s.pth" % mode)
        if not os.path.exists(path):
            print("==========>Train the model from scratch")
            return
        checkpoint = torch.load(path, map_location=self.device, weights_only=False)
        self.model.load_state_dict(checkpoint["model_state_dict"])
        if self.trainer is not None:
            self.trainer._dirty = True           # load_state_dict copied in place into the views of the trainer's flat parameter buffer
            if not getattr(self.args, "pretrained", False):          # the reference reads args.pretrained without defining it (base.py:112)
                print("==========>Load the previous optimizer")
                self._load_optimizer_state_dict(checkpoint["optimizer_state_dict"])
                self.start_epoch = checkpoint["epoch"]
                self.bestAP = checkpoint.get("accuracy", -1.0)
        print("==========>Load the model weight from %s, saved at epoch %d" % (self.dir, checkpoint["epoch"]))

    # ---------------------------------------------------------------------------------------------- results files
    def _xywh2cs(self, x, y, w, h):
        center = np.array([x + w * 0.5, y + h * 0.5], dtype=np.float32)
        if w > self.aspectRatio * h:
            h = w / self.aspectRatio
        elif w < self.aspectRatio * h:
            w = h * self.aspectRatio
        scale = np.array([w / self.pixel_std, h / self.pixel_std], dtype=np.float32)
        if center[0] != -1:
            scale = scale * 1.25
        return center, scale

    def saveKeypoints(self, savePreds, preds, bbox, image_id):
        preds = np.concatenate((preds, np.ones((len(preds), self.numKeypoints, 1))), axis=2)
        for j in range(len(preds)):
            center, scale = self._xywh2cs(float(bbox[j][0]), float(bbox[j][1]), float(bbox[j][2]), float(bbox[j][3]))
            savePreds.append({"category_id": 1, "center": center.tolist(), "image_id": int(image_id[j]), "scale": scale.tolist(), "score": 1.0,
                              "keypoints": preds[j].reshape(self.numKeypoints * 3).tolist()})
        return savePreds

    def writeKeypoints(self, preds):
        with open(os.path.join(self.dir, "test_results.json" if self.args.eval else "val_results.json"), "w") as fp:
            json.dump(preds, fp)

    # ---------------------------------------------------------------------------------------------- loops
    def eval(self, visualization=False, epoch=-1):
        """run.py:35-63.  The per-batch keypoints stay on the device (no synchronising read-back inside the loop); the results file
        the reference writes is produced from ONE transfer after the loop, and AP comes from the device-resident poses
        (``evaluatePoses`` -> hupr_keypoint_oks) — identical to evaluating the file (tests/test_runner_gpu.py)."""
        self.model.eval()
        poses, boxes, ids, losses = [], [], [], []
        for batch in self.testLoader:
            hori = batch["VRDAEmap_hori"].float().to(self.device)
            vert = batch["VRDAEmap_vert"].float().to(self.device)
            preds = self.model(hori, vert)
            loss, loss2, pred2d, _ = self.lossComputer.computeLoss(preds, batch["jointsGroup"], host=False)
            poses.append(pred2d.float() * self.imgHeatmapRatio)
            boxes.append(torch.as_tensor(batch["bbox"]).float())
            ids.extend(int(i) for i in batch["imageId"])
            losses.append(loss.detach().reshape(1) if isinstance(loss, torch.Tensor) else torch.tensor([float(loss)]))
        savePreds = []
        if poses:
            poses_dev = torch.cat(poses)
            self.saveKeypoints(savePreds, poses_dev.cpu().numpy(), torch.cat(boxes).numpy(), ids)
        self.writeKeypoints(savePreds)
        self.last_eval_loss = float(torch.cat([l.float().cpu() for l in losses]).mean()) if losses else float("nan")
        if not poses:
            return float("nan")
        if getattr(self.args, "keypoints", False):
            self.testSet.evaluateEach(self.dir)
        stats = self.testSet.evaluatePoses(ids, poses_dev)
        self.testSet._print_stats(stats)
        return float(stats[0])

    def _loss_weights(self):
        """The (alpha, beta) schedule LossComputer.computeLoss applies per call (misc/losses.py:36-42 of the reference): with
        TRAINING.lossDecay == -1 the objective is loss1 + loss2; otherwise alpha*loss1 + beta*loss2 with alpha ramping up by lossDecay."""
        lc = self.lossComputer
        if lc.alpha < 1.0:
            lc.alpha += lc.lossDecay
            lc.beta -= lc.lossDecay
        return (1.0, 1.0) if lc.lossDecay == -1 else (lc.alpha, lc.beta)

    def train(self):
        for epoch in range(self.start_epoch, self.cfg.TRAINING.epochs):
            self.model.train()
            loss_list = []
            for idxBatch, batch in enumerate(self.trainLoader):
                hori = batch["VRDAEmap_hori"].float().to(self.device)
                vert = batch["VRDAEmap_vert"].float().to(self.device)
                w1, w2 = self._loss_weights()
                loss, loss2 = self.trainer.forward_backward(hori, vert, batch["jointsGroup"], loss_weights=(w1, w2))
                if (w1, w2) != (1.0, 1.0):               # the reported loss is the optimised objective, as in the reference
                    loss = w1 * self.trainer.last_losses[2] + w2 * loss2
                self.trainer.all_reduce_gradients()
                self.trainer.optimizer_step()
                if idxBatch % self.cfg.TRAINING.lrDecayIter == 0:
                    self.adjustLR(epoch)
                loss_list.append(float(loss))
                if not math.isfinite(loss_list[-1]):
                    raise FloatingPointError("non-finite training loss at epoch %d batch %d" % (epoch, idxBatch))
            accAP = self.eval(visualization=False, epoch=epoch)
            self.saveModelWeight(epoch, accAP)
            self.saveLosslist(epoch, loss_list, "train")
