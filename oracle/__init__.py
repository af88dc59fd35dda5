"""CPU oracle for the descriptor-matching hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package (see oracle/oracle.c for the parity statement).
"""
from .oracle import *  # noqa: F401,F403
