"""CPU restatement of the pyfds time-stepping path: TEST INFRASTRUCTURE ONLY (see restate.py).

Nothing under ``pyfds_b200/`` imports this package."""
