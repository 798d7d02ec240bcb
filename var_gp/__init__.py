"""Import-path shim: ``from var_gp.vargp import VARGP`` etc. resolve to the B200-native implementation
in ``vargp_b200`` (drop-in for the reference package of the same name)."""
