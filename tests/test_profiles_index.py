"""profiles/INDEX.md maps the claims of DESIGN.md to evidence files: every file it names plainly must exist."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_file_named_in_the_index_exists():
    txt = open(os.path.join(ROOT, "profiles", "INDEX.md")).read()
    names = set(re.findall(r"`([A-Za-z0-9_.\-]+\.(?:json|txt|log|csv))`", txt))
    assert len(names) > 30
    missing = sorted(n for n in names if not os.path.exists(os.path.join(ROOT, "profiles", n)))
    assert not missing, missing


def test_design_md_cites_existing_profiles():
    txt = open(os.path.join(ROOT, "DESIGN.md")).read()
    names = set(re.findall(r"`(?:profiles/)?(r[12][a-z]_[A-Za-z0-9_.\-]+\.(?:json|txt|log|csv))`", txt))
    missing = sorted(n for n in names if not os.path.exists(os.path.join(ROOT, "profiles", n)))
    assert not missing, missing
