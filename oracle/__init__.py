"""CPU oracle (test infrastructure). See detr_oracle.py / lsap.c headers."""
