"""One eager generator forward between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
from ddgan_b200 import arch
from ddgan_b200.engine import GeneratorEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = arch.make_config()
eng = GeneratorEngine(cfg, B, 'cuda', precision=3)
sd = {k: torch.randn(s) * (0.05 if len(s) > 1 else 0.1) + (1.0 if (len(s) == 1 and k.endswith('weight')) else 0.0) for k, s in eng.shapes.items()}
eng.load_state_dict(sd)
eng.x_in.normal_(); eng.z_in.normal_()
for _ in range(3):
    eng.run_steps()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.run_steps()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done; conv launches:', len([n for n in eng.step_names if n.startswith('conv')]), 'algorithmic conv bytes per forward:', eng.conv_bytes)
