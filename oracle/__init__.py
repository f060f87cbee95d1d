"""CPU oracle of the JAX-in-Cell Boris hot path -- TEST INFRASTRUCTURE, never imported by the product path.

literal.py      expression-level NumPy restatement of the reference (O(N*G)), pinned by the reference's own KATs
closed_form.py  O(N) closed form of the same arithmetic (what the CUDA kernels implement), validated against literal.py
"""
