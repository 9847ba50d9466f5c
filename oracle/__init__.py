"""oracle/ -- TEST INFRASTRUCTURE (CPU restatement of the reference path). See bns_oracle.py."""
