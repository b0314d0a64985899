"""DESIGN.md must not drift from the code: every environment switch the product reads is listed in its switches table,
and the entry-point count it states equals what include/opengenie_b200.h declares."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*parts):
    with open(os.path.join(ROOT, *parts)) as f:
        return f.read()


def test_every_environment_switch_is_documented():
    design = _read('DESIGN.md')
    sources = glob.glob(os.path.join(ROOT, 'open_genie_b200', '**', '*.cu'), recursive=True) + \
        glob.glob(os.path.join(ROOT, 'open_genie_b200', '**', '*.py'), recursive=True) + [os.path.join(ROOT, 'bench.py')]
    used = set()
    for path in sources:
        with open(path) as f:
            text = f.read()
        used |= set(re.findall(r'getenv\("(OG_[A-Z0-9_]+)"\)', text))
        used |= set(re.findall(r"environ\.get\('(OG_[A-Z0-9_]+)'", text))
    assert used, 'no switches found: the patterns of this test are stale'
    missing = sorted(v for v in used if v not in design)
    assert not missing, f'environment switches read by the code but not described in DESIGN.md: {missing}'


def test_entry_point_count_in_design_matches_header():
    header = _read('include', 'opengenie_b200.h')
    declared = re.findall(r'^(?:int|int64_t|size_t|uint64_t|const char\*)\s+(og_[a-z0-9_]+)\(', header, flags=re.M)
    stated = re.search(r'(\d+) entry points', _read('DESIGN.md'))
    assert stated, 'DESIGN.md no longer states the number of entry points'
    assert int(stated.group(1)) == len(set(declared)), (int(stated.group(1)), len(set(declared)))
