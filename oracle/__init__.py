"""CPU oracle (test infrastructure only; see region_oracle.py).  Never imported by the product."""
