"""Empty stand-in so the reference preprocessing package imports (slide_processor.py:14). Test infrastructure only."""
