"""Entry point: host-side mirror of /root/reference/main.py — same flags, same YAML schema, same dispatch.

    python -m hupr_b200.main --config mscsa_prgcn.yaml --dir <run>            # train (resumes ./logs/<run>/checkpoint.pth)
    python -m hupr_b200.main --config mscsa_prgcn.yaml --dir <run> --eval     # evaluate ./logs/<run>/model_best.pth
"""
import argparse
import ast

from .config import load_config
from .tools import Runner


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("--seed", type=int, default=0, metavar="S", help="random seed (default: 0)")
    parser.add_argument("--dir", type=str, default="test", metavar="B", help="directory of saving/loading")
    parser.add_argument("--visDir", type=str, default="none", metavar="B", help="directory of visualization")
    parser.add_argument("--config", type=str, default="mscsa_prgcn.yaml", metavar="B", help="file under ./config")
    parser.add_argument("--gpuIDs", default=[0], type=ast.literal_eval, help="IDs of GPUs to use")
    parser.add_argument("--eval", action="store_true")
    parser.add_argument("-sr", "--sampling_ratio", type=int, default=1, help="sampling ratio for training/test (default: 1)")
    parser.add_argument("--keypoints", action="store_true", help="print out the APs of all keypoints")
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    cfg = load_config("./config/" + args.config)
    runner = Runner(args, cfg)
    if args.eval:
        runner.loadModelWeight("model_best")
        return runner.eval(visualization=False)
    runner.loadModelWeight("checkpoint")
    return runner.train()


if __name__ == "__main__":
    main()
