import sys, torch, numpy as np
sys.path.insert(0, '.')
from tests.helpers import load_case, t
from tests.test_race_gpu import make_env
name = sys.argv[1]
for variant in ('brute', 'fast'):
    c = load_case(name)
    env = make_env(c, variant)
    states, _ = env.reset(t(c['tracks']))
    n = 0
    for s in range(c['actions'].shape[0]):
        fin = env.finished()
        states, rewards = env.step(t(c['actions'][s]).cuda())
        for k, v in (('pos', env.positions), ('dir', env.directions), ('speed', env.speeds), ('alive', env.alive), ('scores', env.scores)):
            ref = torch.from_numpy(c[k][s + 1])
            d = (v.cpu() != ref)
            if d.any() and n < 6:
                n += 1
                idx = d.nonzero()[0].tolist()
                print(variant, 'step', s, k, 'idx', idx, 'got', v.cpu()[tuple(idx)].item(), 'ref', ref[tuple(idx)].item(),
                      'width', states.size(-1), c['widths'][s + 1], 'alive_ref_prev', c['alive'][s][idx[0]].tolist(), 'alive_ref', c['alive'][s+1][idx[0]].tolist(),
                      'steps', env.steps, 'fin', fin, c['finished'][s])
    print(variant, 'mismatch reports', n)
