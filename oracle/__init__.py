"""TEST INFRASTRUCTURE ONLY -- see oracle/README.md.  Never imported by pyhmmer_b200/."""
