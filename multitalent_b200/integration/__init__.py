"""Reference-side glue: files a MultiTalent maintainer drops into a reference checkout."""
