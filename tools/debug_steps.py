"""Run an engine step by step with synchronisation to localise faults / NaNs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'denoising-diffusion-gan_b200'))
import torch
from oracle import ddgan_oracle as O
from ddgan_b200.engine import GeneratorEngine, DiscriminatorEngine, Act

which = sys.argv[1] if len(sys.argv) > 1 else 'gen_tiny'
cfg = O.tiny_config() if which == 'gen_tiny' else O.cifar10_config()
B = 3
eng = GeneratorEngine(cfg, B, 'cuda')
eng.load_state_dict(O.randomize_params(O.ncsnpp_param_shapes(cfg), seed=7))
eng.x_in.normal_(); eng.z_in.normal_(); eng.t_in.copy_(torch.tensor([0, 3, 1]))
torch.cuda.synchronize()

def tensors(obj, seen, out, depth=0):
    if id(obj) in seen or depth > 6:
        return
    seen.add(id(obj))
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda and obj.is_floating_point():
            out.append(obj)
    elif isinstance(obj, Act):
        out.append(obj.buf)
    elif isinstance(obj, (list, tuple)):
        for o in obj: tensors(o, seen, out, depth + 1)
    elif isinstance(obj, dict):
        for o in obj.values(): tensors(o, seen, out, depth + 1)
    elif hasattr(obj, '__dict__') and not isinstance(obj, type):
        for o in vars(obj).values(): tensors(o, seen, out, depth + 1)

# closures hold most buffers: walk closure cells of steps too
allt = []
seen = set()
tensors(eng.__dict__, seen, allt)
for fn in eng.steps:
    if fn.__closure__:
        for c in fn.__closure__:
            try: tensors(c.cell_contents, seen, allt)
            except ValueError: pass
    if fn.__defaults__:
        tensors(list(fn.__defaults__), seen, allt)
print('built', len(eng.steps), 'steps;', len(allt), 'tensors tracked', flush=True)
bad = set()
prev = {}
lo, hi = int(os.environ.get('LO', 1000)), int(os.environ.get('HI', -1))
for t in allt: prev[id(t)] = float(t.abs().max()) if t.numel() else 0.0
for rep in range(2):
    for i, (fn, name) in enumerate(zip(eng.steps, eng.step_names)):
        try:
            fn()
            torch.cuda.synchronize()
        except Exception as e:
            print(f'rep {rep} STEP {i} {name} FAILED: {str(e)[:200]}', flush=True)
            sys.exit(1)
        if rep == 0 and lo - 1 <= i <= hi:
            for t in allt:
                m = float(t.abs().max()) if t.numel() else 0.0
                if prev.get(id(t)) != m:
                    if i >= lo: print(f'   step {i} {name[:60]}: tensor {tuple(t.shape)} {t.dtype} maxabs {prev.get(id(t))} -> {m:.4g}', flush=True)
                    prev[id(t)] = m
        for t in allt:
            if id(t) not in bad and t.dtype != torch.float64 and not bool(torch.isfinite(t).all()):
                bad.add(id(t))
                print(f'rep {rep} after step {i} ({name}): tensor shape {tuple(t.shape)} became non-finite '
                      f'({int((~torch.isfinite(t)).sum())} elements)', flush=True)
    print('rep', rep, 'out finite:', bool(torch.isfinite(eng.out).all()), flush=True)
