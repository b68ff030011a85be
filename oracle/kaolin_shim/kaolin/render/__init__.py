"""Only present so `import kaolin as kal` + attribute access in deprecated reference code parses."""
