"""Stand-in for the `gensim` import of the reference's data_loader/dataset.py:3 (TEST INFRASTRUCTURE ONLY; gensim is not installed
offline).  KeyedVectors is only consulted for test_topk != -1 (dataset.py:326-329), which the parity probes do not use."""
