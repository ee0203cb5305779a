"""Analysis utilities (the role of the reference's ``notebooks/01..16``): schedules, rank / singular-value analysis of the
learned updates, dataset and checkpoint checks.  Every tool is a small CLI (``python -m tools.<name> --help``) whose core is an
importable function so it can be tested on CPU."""
