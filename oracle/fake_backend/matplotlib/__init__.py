"""Fake ``matplotlib`` -- test infrastructure; plotting is a no-op."""
