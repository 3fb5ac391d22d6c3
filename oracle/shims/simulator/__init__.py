"""TEST INFRASTRUCTURE — stand-in package for the absent third-party `simulator` (see game/connect.py)."""
