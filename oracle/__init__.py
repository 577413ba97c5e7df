"""ORACLE — test infrastructure only. CPU restatement of the reference's solver + integrator path (see oracle/bepu_math.h).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package."""
