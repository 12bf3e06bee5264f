"""TEST INFRASTRUCTURE ONLY: CPU checker for the CUDA k-mer counting path (see kmer_oracle.h)."""
