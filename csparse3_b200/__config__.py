"""Backend switch, same name as the reference's (src/CSparse3/__config__.py:1).  In this package the flat
kernels always come from the B200 backend (csc_b200); the flag is kept so code written against the
reference (`__config__.NATIVE = True`, src/test/test1_operations.py:8) keeps working."""
NATIVE = True
