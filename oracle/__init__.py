"""TEST INFRASTRUCTURE ONLY: CPU oracle (C restatement + reference harness).

Nothing under the product package imports this.  See oracle/swimmer_oracle.c."""
