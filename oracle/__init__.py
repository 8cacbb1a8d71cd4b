"""CPU oracle (test infrastructure only -- see oracle/exact_search.py header)."""
