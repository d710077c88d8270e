"""Module-tree mirrors of the reference networks (see networks.py)."""
