import sys, os, time
sys.path.insert(0, '/root/repo' if os.path.isdir('/root/repo/f2cnn_b200') else '.')
import torch
from f2cnn_b200 import cnn
m = cnn.seeded_model(0)
x = torch.rand(46240, 11, 128, device='cuda')
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/reps*1e3
print('fp32 default', timed(lambda: cnn.predict(m, x)))
torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
print('tf32 allowed', timed(lambda: cnn.predict(m, x)))
print('bf16 autocast', timed(lambda: cnn.predict(m, x, autocast_dtype=torch.bfloat16)))
torch.backends.cudnn.benchmark = True
print('bf16 + cudnn.benchmark', timed(lambda: cnn.predict(m, x, autocast_dtype=torch.bfloat16)))
print('fp32 + cudnn.benchmark', timed(lambda: cnn.predict(m, x)))
m2 = cnn.seeded_model(0).to(memory_format=torch.channels_last)
def cl(dt):
    out=[]
    with torch.no_grad(), torch.autocast('cuda', dtype=dt):
        for i in range(0, x.shape[0], 8192):
            xx = x[i:i+8192].unsqueeze(1).contiguous(memory_format=torch.channels_last)
            import torch.nn.functional as F
            h = F.relu(m2.c1(xx)); h = F.max_pool2d(F.relu(m2.c2(h)), 2); h = F.relu(m2.c3(h)); h = F.max_pool2d(F.relu(m2.c4(h)), 2)
            h = h.permute(0,2,3,1).reshape(h.shape[0], -1); out.append(F.softmax(m2.d2(F.relu(m2.d1(h))), dim=1))
    return torch.cat(out)
print('channels_last bf16', timed(lambda: cl(torch.bfloat16)))
print('channels_last fp16', timed(lambda: cl(torch.float16)))
