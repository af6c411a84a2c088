"""CPU oracle for the annembed hot path -- TEST INFRASTRUCTURE ONLY (see annembed_oracle.c)."""
