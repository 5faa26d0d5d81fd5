"""Make baseline/_ref: the UNMODIFIED reference packages bench.py's CPU arm can time (`kind: "reference"`).

    python tools/make_baseline_ref.py [/root/reference]

Copies the reference's `games/` and `utils/` packages (the only ones `games/race.py` imports) and `agents/`, `policies/`
(the stock PPOAgent / LSTMPolicy that tests/test_rollout_graph_gpu.py drives through GraphedRollout) into baseline/_ref/,
which is git-ignored (it never enters this repository's history) but travels to the GPU box with the working tree.
The reference has no setup.py / pyproject, so `pip install --target baseline/_ref /root/reference` fails; this is
the same outcome by plain copy.  Without baseline/_ref bench.py times the torch-op restatement instead
(`kind: "port"`).  The reference's C++ helper (games/game_helpers.cpp) needs Boost.Geometry, which this image
lacks: importing `games` tries to build it, fails, and the reference itself falls back to its torch path
(games/race.py:89-101) - that path, on device=cpu, is what gets timed.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    dst = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(src, 'games')):
        sys.exit('no reference at %s' % src)
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    for pkg in ('games', 'utils', 'agents', 'policies'):      # agents / policies: tests/test_rollout_graph_gpu.py (PPOAgent.act)
        shutil.copytree(os.path.join(src, pkg), os.path.join(dst, pkg),
                        ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    if os.path.isdir(os.path.join(src, 'learned')):         # the shipped agents (train-gan.py:22, 46-50): config 3 of bench.py
        shutil.copytree(os.path.join(src, 'learned'), os.path.join(dst, 'learned'))
    print('copied %s/{games,utils,agents,policies,learned} -> %s' % (src, dst))


if __name__ == '__main__':
    main()
