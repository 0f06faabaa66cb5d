"""TEST INFRASTRUCTURE. CPU oracle for the hot path; see oracle/oracle.cpp header.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
