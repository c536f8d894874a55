"""torchvision ResNet-50, random weights with trained-looking BatchNorm statistics, as TorchScript `model.pt` (the file the
reference places for Triton's libtorch backend).  `--arch resnet18` for a small one."""
import os
import sys

import torch
import torchvision

args = [a for a in sys.argv[1:] if not a.startswith("--")]
arch = sys.argv[sys.argv.index("--arch") + 1] if "--arch" in sys.argv else "resnet50"
out = args[0] if args else "."
if os.path.isdir(out) or not out.endswith(".pt"):
    os.makedirs(out, exist_ok=True)
    out = os.path.join(out, "resnet_model.pt")
torch.manual_seed(0)
m = getattr(torchvision.models, arch)(weights=None).eval()
with torch.no_grad():
    for name, mod in m.named_modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.uniform_(0.75, 1.25).mul_(0.3 if name.endswith("bn3") else 1.0)
            mod.bias.normal_(0, 0.1)
torch.jit.script(m).save(out)
print(out)
