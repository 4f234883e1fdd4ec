import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from hupr_b200.preprocessing.process_iwr1843 import cascade_i16, FRAME_WORDS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
adc = torch.randint(-2048, 2048, (n, FRAME_WORDS), dtype=torch.int16, device='cuda')
out = torch.empty((n, 16, 64, 64, 8), dtype=torch.complex64, device='cuda')
for _ in range(3): cascade_i16(adc, out)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
for i in range(10):
    ev[i].record(); cascade_i16(adc, out)
ev[10].record(); torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i+1]) for i in range(10)]
t = np.median(ms) * 1e-3
byt = n * (786432 + 4194304)
print("n=%d median %.3f ms  %.1f k frame-sensors/s  %.1f GB/s (%.1f%% of 6539.9)" % (n, t*1e3, n/t/1e3, byt/t/1e9, 100*byt/t/1e9/6539.9))
print("all ms:", ["%.3f" % m for m in ms])
