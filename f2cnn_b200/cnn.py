"""Forward pass of the reference's F2-direction CNN (scripts/CNN/Training.py:93-114) in PyTorch,
for the `cnn eval*` path: Conv32 3x3 same -> Conv32 3x3 -> MaxPool2 -> Conv64 3x3 same ->
Conv64 3x3 -> MaxPool2 -> Flatten(1920) -> Dense516 -> Dense2 softmax, ReLU after every conv /
the first dense, dropout inactive at inference.

This is a consumer of the hot path, not part of it: the
convolutions go through cuDNN as plain library calls.  No trained model ships with the
reference and Keras is not installed here, so weights are either seeded (benchmarks, tests) or
loaded from arrays in Keras layout with `load_keras_arrays`."""
import torch
import torch.nn.functional as F


class F2CNN(torch.nn.Module):
    def __init__(self, dots=11, channels=128, num_classes=2):
        super().__init__()
        self.c1 = torch.nn.Conv2d(1, 32, 3, padding=1)
        self.c2 = torch.nn.Conv2d(32, 32, 3)
        self.c3 = torch.nn.Conv2d(32, 64, 3, padding=1)
        self.c4 = torch.nn.Conv2d(64, 64, 3)
        h = ((dots - 2) // 2 - 2) // 2
        w = ((channels - 2) // 2 - 2) // 2
        self.flat = 64 * h * w  # 1920 for 11 x 128
        self.d1 = torch.nn.Linear(self.flat, 516)
        self.d2 = torch.nn.Linear(516, num_classes)

    def forward(self, frames):
        """frames: (N, dots, channels) -> (N, num_classes) softmax scores [falling, rising]."""
        x = frames.unsqueeze(1)
        x = F.relu(self.c1(x))
        x = F.max_pool2d(F.relu(self.c2(x)), 2)
        x = F.relu(self.c3(x))
        x = F.max_pool2d(F.relu(self.c4(x)), 2)
        # Keras flattens channels-last (H, W, C); keep that order so Keras dense weights fit
        x = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)
        x = F.relu(self.d1(x))
        return F.softmax(self.d2(x), dim=1)

    @torch.no_grad()
    def load_keras_arrays(self, arrays):
        """arrays: [k1, b1, ..., k4, b4, W1, c1, W2, c2] in Keras layout (conv kernels HWIO,
        dense kernels (in, out))."""
        convs = [self.c1, self.c2, self.c3, self.c4]
        for i, conv in enumerate(convs):
            conv.weight.copy_(torch.as_tensor(arrays[2 * i]).permute(3, 2, 0, 1))
            conv.bias.copy_(torch.as_tensor(arrays[2 * i + 1]))
        for j, lin in enumerate([self.d1, self.d2]):
            lin.weight.copy_(torch.as_tensor(arrays[8 + 2 * j]).t())
            lin.bias.copy_(torch.as_tensor(arrays[9 + 2 * j]))


def seeded_model(seed=0, dots=11, channels=128, device="cuda", dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    m = F2CNN(dots, channels)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.05 if p.dim() > 1 else 0.01))
    return m.to(device=device, dtype=dtype).eval()


@torch.no_grad()
def predict(model, frames_dev, batch=8192, autocast_dtype=None, channels_last=False):
    """Scores for all frames, in batches (46 240 frames per 3 s utterance).  autocast_dtype =
    torch.bfloat16 runs the convolutions and dense layers on the tensor cores in bf16 (inputs are
    normalised to [0, 1], scores differ by ~1e-2); None keeps the parameters' precision.
    channels_last=True lays the convolution weights and activations out NHWC, the layout cuDNN's
    tensor-core kernels want (57 ms fp32 -> 49 ms bf16 -> 33 ms bf16 NHWC per 3 s utterance); it
    converts the module's parameters in place, the results do not depend on it."""
    out = []
    if channels_last:
        model.to(memory_format=torch.channels_last)
    dt = next(model.parameters()).dtype
    for i in range(0, frames_dev.shape[0], batch):
        x = frames_dev[i:i + batch].to(dt)
        if autocast_dtype is not None:
            with torch.autocast("cuda", dtype=autocast_dtype):
                y = model(x)
            out.append(y.float())
        else:
            out.append(model(x))
    return torch.cat(out) if out else torch.zeros((0, 2), device=frames_dev.device)
