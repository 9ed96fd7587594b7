"""jwave-pro_b200: B200-native MODWT / FWT / WPT behind JWave-Pro's transform API.

Import it as `jwave_pro_b200` (the shim package next to this directory)."""
