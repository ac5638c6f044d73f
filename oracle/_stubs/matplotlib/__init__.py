"""Empty stand-in so that the read-only reference module (which imports
matplotlib.pyplot at module scope) can be imported in the build container.
Used ONLY by oracle/gen_golden.py; never by the product."""
