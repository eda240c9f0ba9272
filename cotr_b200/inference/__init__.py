"""Mirror of the reference's COTR/inference package (zoom-in loop host code)."""
