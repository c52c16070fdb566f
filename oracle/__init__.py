"""TEST INFRASTRUCTURE ONLY: CPU oracle of the denoiser hot path (see oracle/egnn_oracle.py).
May be imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference legs."""
