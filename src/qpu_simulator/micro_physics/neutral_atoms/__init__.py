"""see qpu_simulator/__init__.py"""
