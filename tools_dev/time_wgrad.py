"""Time hupr_conv_wgrad against the position-major path (to_kmajor + split-K GEMMs) on the training step's shapes (batch 16)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hupr_b200 import ops
from hupr_b200.ops import SplitTensor

B = int(os.environ.get("WG_BATCH", "16"))
SHAPES = [   # name, cin, cout, d, h, w, kernel, pad
    ("l1.conv0 32->64", 64, 64, 8, 64, 64, (3, 3, 3), (1, 1, 1)),
    ("l1.conv1 64->128", 64, 128, 8, 64, 64, (3, 3, 3), (1, 1, 1)),
    ("l1.conv2 64->64", 64, 64, 8, 64, 64, (3, 3, 3), (1, 1, 1)),
    ("l2.conv1 64->256", 64, 256, 4, 32, 32, (3, 3, 3), (1, 1, 1)),
    ("l2.conv2 128->128", 128, 128, 4, 32, 32, (3, 3, 3), (1, 1, 1)),
    ("l2.conv1 128->256", 128, 256, 4, 32, 32, (3, 3, 3), (1, 1, 1)),
    ("l3.conv1 128->512", 128, 512, 2, 16, 16, (3, 3, 3), (1, 1, 1)),
    ("l3.conv2 256->256", 256, 256, 2, 16, 16, (3, 3, 3), (1, 1, 1)),
    ("l3.conv1 256->512", 256, 512, 2, 16, 16, (3, 3, 3), (1, 1, 1)),
    ("merge1 8x1x1 64", 64, 64, 8, 64, 64, (8, 1, 1), (0, 0, 0)),
    ("proj1 64->256", 64, 256, 1, 64, 64, (1, 1, 1), (0, 0, 0)),
    ("proj2 128->512", 128, 512, 1, 32, 32, (1, 1, 1), (0, 0, 0)),
    ("dec3 1024->512", 1024, 512, 1, 16, 16, (1, 3, 3), (0, 1, 1)),
    ("dec2 640->256", 640, 256, 1, 32, 32, (1, 3, 3), (0, 1, 1)),
    ("dec1 320->128", 320, 128, 1, 64, 64, (1, 3, 3), (0, 1, 1)),
    ("dec1b 64->64", 64, 64, 1, 64, 64, (1, 3, 3), (0, 1, 1)),
]


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


rows = []
for name, cin, cout, d, h, w, kernel, pad in SHAPES:
    torch.manual_seed(0)
    d_out = d + 2 * pad[0] - kernel[0] + 1
    X = SplitTensor.from_float(torch.randn(B, d, h, w, cin, device="cuda"))
    DY = SplitTensor.from_float(torch.randn(B, d_out, h, w, cout, device="cuda"))
    taps = kernel[0] * kernel[1] * kernel[2]
    out = torch.zeros((taps, cin, cout), device="cuda")
    t_new = timeit(lambda: ops.conv_wgrad_direct(X, 0, cin, DY, 0, cout, kernel, pad, out=out))
    # position-major path as training.ConvOp.wgrad runs it
    geom = ops.KMajorGeometry(B, d, h, w, pad)
    rows_ = -(-cout // 128) * 128
    xt = SplitTensor.empty((kernel[2] * cin, geom.ppad), "cuda", zero=True)
    dyt = SplitTensor.empty((rows_, geom.ppad), "cuda", zero=True)
    acc = torch.zeros((kernel[0] * kernel[1], rows_, kernel[2] * cin), device="cuda")

    def old():
        ops.to_kmajor_multi(X, 0, cin, geom, xt, -pad[2], kernel[2], cin)
        ops.to_kmajor(DY, 0, cout, geom, dyt)
        ops.conv_wgrad(xt, cin, dyt, rows_, geom, kernel, acc)
    t_old = timeit(old, reps=3)
    fl = 2.0 * B * d_out * h * w * cin * cout * taps
    # parity between the two paths (same inputs)
    out.zero_(); acc.zero_()
    ops.conv_wgrad_direct(X, 0, cin, DY, 0, cout, kernel, pad, out=out); old(); torch.cuda.synchronize()
    ref = acc.view(kernel[0] * kernel[1], rows_, kernel[2], cin).permute(0, 2, 1, 3).reshape(taps, rows_, cin)[:, :cout].permute(0, 2, 1)
    err = float((out - ref).abs().max() / ref.abs().max())
    rows.append({"shape": name, "new_ms": round(t_new, 4), "old_ms": round(t_old, 4), "new_tflops": round(fl / t_new / 1e9, 1),
                 "old_tflops": round(fl / t_old / 1e9, 1), "rel_diff": err})
    print(json.dumps(rows[-1]), flush=True)
    del X, DY, xt, dyt, acc, out
print("TOTAL new %.3f ms, old %.3f ms" % (sum(r["new_ms"] for r in rows), sum(r["old_ms"] for r in rows)))
