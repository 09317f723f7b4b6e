"""repeat_expand_2d on the device: bit-identical to the reference-style per-column walk run on a CUDA tensor, and what each costs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ns2vc_b200.frontend import repeat_expand_2d


def walk(content, target_len):                      # utils.py:482-496 as the reference runs it (content on the GPU)
    src_len = content.shape[-1]
    target = torch.zeros([content.shape[0], target_len], dtype=torch.float).to(content.device)
    temp = torch.arange(src_len + 1) * target_len / src_len
    current_pos = 0
    for i in range(target_len):
        if i < temp[current_pos + 1]:
            target[:, i] = content[:, current_pos]
        else:
            current_pos += 1
            target[:, i] = content[:, current_pos]
    return target


def timed(fn, n):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) * 1e3 / n

for src, tgt in ((530, 1000), (160, 300), (1100, 2048)):
    c = torch.randn((256, src), generator=torch.Generator().manual_seed(src)).cuda()
    a, ta = timed(lambda: walk(c, tgt), 3)
    b, tb = timed(lambda: repeat_expand_2d(c, tgt), 20)
    print(f"t_src={src} -> {tgt} frames: reference-style walk on the device {ta:.2f} ms, host walk + one gather {tb:.3f} ms, equal={torch.equal(a, b)}")
