"""Empty stand-in for `ipdb` (imported but unused at reference model/model.py:8).

TEST INFRASTRUCTURE ONLY. Lives on PYTHONPATH only while the unmodified
reference model files are executed to generate / check golden vectors.
"""
