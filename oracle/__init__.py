"""TEST INFRASTRUCTURE: loaders for the CPU oracles.  Importable only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from convectionkernels_b200."""
