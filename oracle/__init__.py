"""CPU oracle for the Hex8 CVFEM assembly path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this
package; the product (nalu_b200/) never does.  See nalu_oracle.h for the parity status.
"""
from .oracle import *  # noqa: F401,F403
