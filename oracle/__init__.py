"""ORACLE — test infrastructure only.

CPU restatement of the reference's hot path (rakhimovv/lvt, package `vidgen`), used as the
parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
arm.  Nothing under lvt_b200/ imports this package; the product path fails loudly without its
CUDA library instead of falling back to this code.

  vq_oracle.c / vq.py : plain-C codebook search with ATen's exact fp32 operation order
  lvt_oracle.py       : torch-CPU fp32 functional restatement of the VQ-VAE and DSFVT modules
  ref_shim.py         : in-memory stand-ins for fvcore/yacs/termcolor so that the UNMODIFIED
                        reference can be imported from /root/reference (authoring container
                        only) to pin this oracle and to generate tests/golden/*
"""
