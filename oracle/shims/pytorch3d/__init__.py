"""TEST INFRASTRUCTURE ONLY -- empty stand-in for pytorch3d (only used by an uncalled reference helper, dvmvs/utils.py:157-202)."""
