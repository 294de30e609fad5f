"""Vendor the UNMODIFIED reference subset that the hot path needs into baseline/_ref/ (git-ignored, NOT
gpurun-ignored: it travels to the GPU box with the snapshot) -- TEST / BENCHMARK INFRASTRUCTURE ONLY.

    python oracle/vendor_reference.py            # no-op when /root/reference is absent (GPU box)

The files are copied byte for byte from /root/reference/Construction_based (nets/, problems/, utils/, train.py,
eval.py, reinforce_baselines.py, options.py, generate_data.py, agh_baseline.py, the shipped agh20 validation set and
args.json files); nothing is edited and nothing lands in the git history.  `oracle/reference_arm.py` imports the copy
(with empty stand-ins for the missing matplotlib / tensorboard_logger modules) as bench.py's `--impl reference` arm
and as an extra check of the oracle.  A sha256 manifest is written next to the copy.
"""
import hashlib
import json
import os
import shutil
import sys

SRC = '/root/reference/Construction_based'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, 'baseline', '_ref')
DIRS = ['nets', 'problems', 'utils']
FILES = ['train.py', 'eval.py', 'reinforce_baselines.py', 'options.py', 'generate_data.py', 'agh_baseline.py', 'run.py',
         'data/agh/agh20_validation_seed4321.pkl', 'data/20/args.json', 'data/50/args.json', 'data/100/args.json']


def vendor(verbose=True) -> bool:
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    manifest = {}
    for d in DIRS:
        shutil.copytree(os.path.join(SRC, d), os.path.join(DST, d), dirs_exist_ok=True,
                        ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    for f in FILES:
        os.makedirs(os.path.dirname(os.path.join(DST, f)), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    for base, _, names in os.walk(DST):
        if '__pycache__' in base:
            continue
        for nm in sorted(names):
            if nm == 'MANIFEST.json' or nm.endswith('.pyc'):
                continue
            p = os.path.join(base, nm)
            os.chmod(p, 0o644)
            manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, 'rb').read()).hexdigest()
    json.dump({'source': SRC, 'files': manifest}, open(os.path.join(DST, 'MANIFEST.json'), 'w'), indent=1, sort_keys=True)
    if verbose:
        print('vendored %d reference files into %s' % (len(manifest), DST))
    return True


if __name__ == '__main__':
    sys.exit(0 if vendor() else 1)
