"""zkb — B200-native BN254 Groth16 proving backend behind eigen-zeth's ProverService."""
