"""Stand-in for the handful of Raysect declarations integration/cherab_b200_shim uses (Raysect 0.8.1 is not installable here)."""
