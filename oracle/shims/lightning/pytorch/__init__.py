"""TEST INFRASTRUCTURE — see ../__init__.py."""
