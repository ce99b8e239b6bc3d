"""Import shims for the reference's missing `helpers` git submodule (jramapuram/helpers@3b7b824, absent from
/root/reference).  Only what /root/reference/main.py touches; behaviour inferred from the call sites listed in
SURVEY.md §8(c).  Used by the oracle harness and by INTEGRATION.md's drop-in recipe — never by byol_b200 itself."""
