"""Python-3 restatements of the parts of the reference's analysis tools that define acceptance numbers for the
hot path (growth-rate fit, energy peak, analytic dispersion root).  Test / analysis infrastructure."""
