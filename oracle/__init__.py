"""TEST INFRASTRUCTURE: CPU restatements (oracles) of the libbpk entry points and of the reference PCA sweep."""
