"""Test oracle for skeletor_b200 (CPU restatement + compiled reference loader).

TEST INFRASTRUCTURE.  Nothing in skeletor_b200/ imports this package.
"""
