"""oracle/ — TEST INFRASTRUCTURE ONLY (see magicodec_oracle.py header). Parity unpinned by the reference."""
