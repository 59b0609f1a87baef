"""TEST INFRASTRUCTURE ONLY — CPU oracle of SPEC-PROVISIONAL.md.  PARITY UNPINNED (no reference
source exists to pin it against).  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
