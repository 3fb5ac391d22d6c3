"""TEST INFRASTRUCTURE — see connect.py."""
