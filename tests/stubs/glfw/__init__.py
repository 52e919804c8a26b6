"""import-only stub (test infrastructure): the real package is not installed and is not on the step path."""
