"""Stages the UNMODIFIED reference (spencerkent/vision-transform-codes) into the git-ignored oracle/_ref/.

  python tools/stage_reference.py [--source /root/reference] [--check]

The GPU box has no /root/reference; oracle/_ref/ is ignored by git (no reference source enters the history) but not by
gpurun, so it travels with the snapshot like the built .so does. What is staged is the reference's own Python package,
byte for byte (a manifest of sha256 sums is written next to it and verified by --check and by the tests):

  oracle/_ref/vision_transform_codes/{analysis_transforms,dict_update_rules,training,utils}/**.py

Consumers (test infrastructure and the CPU baseline only -- nothing under vision_transform_codes_b200/ reads it):
  * tests/test_gpu_reference_trainer.py runs the reference's train_dictionary on top of the CUDA drop-ins
  * bench.py --impl reference / cpu_baseline time the reference's own CPU implementation (kind "reference")
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, 'oracle', '_ref')
PACKAGE = 'vision_transform_codes'
SUBTREES = ('analysis_transforms', 'dict_update_rules', 'training', 'utils')


def sha256(path):
  h = hashlib.sha256()
  with open(path, 'rb') as f:
    h.update(f.read())
  return h.hexdigest()


def staged_root():
  """oracle/_ref/vision_transform_codes (the directory the reference puts on sys.path), or None if not staged."""
  root = os.path.join(DEST, PACKAGE)
  return root if os.path.isfile(os.path.join(DEST, 'MANIFEST.json')) else None


def stage(source):
  src_pkg = os.path.join(source, PACKAGE)
  if not os.path.isdir(src_pkg):
    raise SystemExit('no reference package at %s' % src_pkg)
  if os.path.isdir(DEST):
    shutil.rmtree(DEST)
  manifest = {}
  for sub in SUBTREES:
    for dirpath, _, files in os.walk(os.path.join(src_pkg, sub)):
      for name in sorted(files):
        if not name.endswith('.py'):
          continue
        src = os.path.join(dirpath, name)
        rel = os.path.relpath(src, source)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = sha256(dst)
  for name in ('LICENSE',):
    if os.path.isfile(os.path.join(source, name)):
      shutil.copyfile(os.path.join(source, name), os.path.join(DEST, name))
  with open(os.path.join(DEST, 'MANIFEST.json'), 'w') as f:
    json.dump({'source': source, 'files': manifest}, f, indent=1, sort_keys=True)
  return manifest


def check():
  """True when every staged file still has the sha256 recorded at staging time (i.e. nobody edited the copy)."""
  path = os.path.join(DEST, 'MANIFEST.json')
  if not os.path.isfile(path):
    return False
  files = json.load(open(path))['files']
  return all(os.path.isfile(os.path.join(DEST, rel)) and sha256(os.path.join(DEST, rel)) == digest
             for rel, digest in files.items())


if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--source', default='/root/reference')
  ap.add_argument('--check', action='store_true')
  args = ap.parse_args()
  if args.check:
    ok = check()
    print('oracle/_ref %s' % ('matches its manifest' if ok else 'is missing or was edited'))
    sys.exit(0 if ok else 1)
  m = stage(args.source)
  print('staged %d reference files into %s' % (len(m), DEST))
