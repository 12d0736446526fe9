"""Host-side profile (cProfile) of BASELINE cfg 3 / cfg 4 through evaluate(): where the Python time of the lock-step loop goes.
    python tools/profile_cfg.py cfg4 > gpurun_out/cfg4_profile.txt"""
import cProfile, io, os, pstats, sys, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from evreal_b200 import evaluate as ev, synthetic

tag = sys.argv[1] if len(sys.argv) > 1 else 'cfg4'
root = os.path.join(tempfile.gettempdir(), 'evk_profile_' + tag)
shutil.rmtree(root, ignore_errors=True)
spec, dur = bench.write_plugin_tree(root, tag, 0, 1, lambda: torch.cuda.synchronize())
lpw = bench.seeded_lpips_weights('alex') if 'lpips' in spec['metrics'] else None
kw = dict(config_root=os.path.join(root, 'config'), write_files=False, rank=0, world_size=1, lockstep=spec['n_seq'], lpips_weights=lpw)
ev.evaluate([spec['method']], ['std'], [spec['dataset']], spec['metrics'], **kw)        # warm-up
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
ev.evaluate([spec['method']], ['std'], [spec['dataset']], spec['metrics'], **kw)
torch.cuda.synchronize()
pr.disable()
print(dict(ev.last_timings))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue())
