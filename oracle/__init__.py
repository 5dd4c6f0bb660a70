"""ORACLE — test infrastructure only.  See oracle/dcol_oracle.hpp."""
